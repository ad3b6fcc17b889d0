set -x
python -m pytest tests/test_gpu_conv.py tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for c in 0 1; do
export B2D_CONV_CONTIG=$c
python tools/profile_conv.py 2d 88 64 64 64 0 2 50 1
python tools/profile_conv.py 2d 88 128 128 32 0 2 50 1
python tools/profile_conv.py 2d 88 256 256 16 0 2 50 1
done
unset B2D_CONV_CONTIG
python tools/diag_chain.py 88 64 10 | head -3
B2D_CONV_CONTIG=0 python tools/diag_chain.py 88 64 10 | head -3
