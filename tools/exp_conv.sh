set -x
python -m pytest tests/test_gpu_fused_gn.py tests/test_gpu_chain.py -x -q 2>&1 | tail -5
python -m pytest tests/test_gpu_models.py -x -q 2>&1 | tail -3
python tools/profile_unet.py 88 10 2>&1 | tail -4
B2D_UNET_FUSE_GN=0 python tools/profile_unet.py 88 10 2>&1 | tail -4
