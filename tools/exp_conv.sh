set -x
python -m pytest tests/test_gpu_conv.py tests/test_gpu_chain.py -x -q 2>&1 | tail -3
python tools/profile_conv.py 1x1 88 256 768 16 0 2 50
python tools/profile_conv.py 1x1 88 256 256 16 0 2 50
python tools/profile_conv.py 1x1 88 512 1536 8 0 2 50
python tools/profile_conv.py 1x1 88 1024 3072 4 0 2 50
python tools/profile_unet.py 88 10 2>&1 | tail -3
python -m pytest tests/test_gpu_models.py -x -q 2>&1 | tail -3
