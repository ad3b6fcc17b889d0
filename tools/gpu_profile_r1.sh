# Round-1 profile pass (run under gpurun): plain runs first, then the ncu passes of the same commands.
set -x
python tools/profile_conv.py 3d 8 512 256 128 0 2 3 1 > gpurun_out/conv_dom_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_v2 -c 1 -s 3 -o gpurun_out/r1_conv_v2_dominant python tools/profile_conv.py 3d 8 512 256 128 0 2 3 1 > gpurun_out/ncu_dom.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_for_ncu.json 2> gpurun_out/bench_for_ncu.err && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -8
