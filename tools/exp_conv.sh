set -x
python tools/profile_conv.py 1x1 88 256 768 16 0 2 50
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_v2 -s 2 -c 1 -o gpurun_out/inproj2 -f python tools/profile_conv.py 1x1 88 256 768 16 0 2 3 > gpurun_out/ncu_inproj.log 2>&1
tail -2 gpurun_out/ncu_inproj.log
