set -x
python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for st in 1 0; do
python tools/profile_conv.py 2d 88 64 64 64 0 2 50 $st
python tools/profile_conv.py 2d 88 128 128 32 0 2 50 $st
python tools/profile_conv.py 2d 88 256 256 16 0 2 50 $st
python tools/profile_conv.py 2d 88 512 512 8 0 2 50 $st
python tools/profile_conv.py 2d 88 1024 1024 4 0 2 50 $st
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_v2 -s 2 -c 1 -o gpurun_out/l0conv -f python tools/profile_conv.py 2d 88 64 64 64 0 2 3 1 > gpurun_out/ncu_l0.log 2>&1
tail -2 gpurun_out/ncu_l0.log
