for n in 22 88 352; do python tools/profile_conv.py 2d $n 64 64 64 0 2 20 1; python tools/profile_conv.py 2d $n 64 64 64 0 2 20 0; done
python tools/profile_conv.py 2d 88 64 64 64 0 1 20 1
python tools/profile_conv.py 2d 88 64 64 64 0 1 20 0
python tools/profile_conv.py 2d 88 128 128 32 0 2 20 1
python tools/profile_conv.py 2d 88 128 128 32 0 2 20 0
python tools/profile_conv.py 2d 88 1024 1024 4 0 2 20 1
python tools/profile_conv.py 2d 88 1024 1024 4 0 2 20 0
python tools/profile_conv.py 3d 8 128 128 256 0 2 3 1
python tools/profile_conv.py 3d 8 128 128 256 0 2 3 0
python tools/profile_conv.py 3d 8 256 256 128 128 2 3 0
python tools/profile_conv.py 3d 8 256 256 128 256 2 3 0
python tools/profile_conv.py 3d 8 512 512 64 128 2 3 0
python tools/profile_conv.py 3d 8 512 512 64 256 2 3 0
