set -x
python tools/profile_conv.py 3d 8 512 512 64 0 2 3 1 > gpurun_out/conv_dom2_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_v2 -c 1 -s 3 -o gpurun_out/r1_conv_v2_512_dominant python tools/profile_conv.py 3d 8 512 512 64 0 2 3 1 > gpurun_out/ncu_dom2.log 2>&1
python tools/profile_conv.py 3d 8 128 128 256 0 2 3 1 > gpurun_out/conv_128_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_v2 -c 1 -s 3 -o gpurun_out/r1_conv_v2_128 python tools/profile_conv.py 3d 8 128 128 256 0 2 3 1 > gpurun_out/ncu_128.log 2>&1
python tools/profile_unet.py 88 1 > gpurun_out/unet_plain2.txt 2>&1 && \
ncu --set full --clock-control none -k regex:attention_tc -c 1 -s 6 -o gpurun_out/r1_attention_tc python tools/profile_unet.py 88 1 > gpurun_out/ncu_attn.log 2>&1
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r1_launches_unet_step_warm.csv python tools/profile_unet.py 88 1 > gpurun_out/ncu_unet2.log 2>&1
ls gpurun_out/*.ncu-rep
