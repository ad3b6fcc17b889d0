# End-of-round capture (run under gpurun): tests, bench (with the CPU baseline), reference arm, stage profile, then the
# ncu launch list of the same bench command.
set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r1_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
python tools/profile_stages.py 8 3 > gpurun_out/r1_stages_eventtimes_final.txt 2>&1
python tools/profile_scheduler.py > gpurun_out/r1_scheduler.txt 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_for_ncu.json 2> gpurun_out/bench_for_ncu.err && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/profile_unet.py 88 5 > gpurun_out/r1_unet_step_eventtimes_final.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r1_launches_unet_step_warm.csv python tools/profile_unet.py 88 1 nograph > gpurun_out/ncu_unet.log 2>&1
ls -la gpurun_out | tail -5
