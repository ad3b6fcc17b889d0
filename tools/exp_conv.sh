set -x
python tools/profile_conv.py convT 88 128 64 32 0 2 50
python tools/profile_conv.py convT 88 128 64 32 0 2 50 0
python tools/profile_conv.py convT 88 256 128 16 0 2 50
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_v2 -s 2 -c 1 -o gpurun_out/convt -f python tools/profile_conv.py convT 88 128 64 32 0 2 3 > gpurun_out/ncu_convt.log 2>&1
tail -2 gpurun_out/ncu_convt.log
