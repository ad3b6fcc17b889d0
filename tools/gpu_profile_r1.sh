set -x
python tools/profile_stages.py 8 3 > gpurun_out/stages_r1.txt 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/profile_conv.py 8 128 128 256 0 3 > gpurun_out/conv_alone.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 1 -s 2 -o gpurun_out/conv128_r1 python tools/profile_conv.py 8 128 128 256 0 1 > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out
