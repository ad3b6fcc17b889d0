python tools/profile_conv.py 2d 1 64 64 16 0 2 50 1
python tools/profile_conv.py 2d 1 64 64 16 0 2 50 0
python tools/profile_conv.py 2d 1 64 64 16 0 1 50 1
python tools/profile_conv.py 2d 8 64 64 64 0 2 50 1
python tools/profile_conv.py 2d 16 1024 1024 4 0 2 50 1
python tools/profile_conv.py 2d 88 256 256 16 0 2 50 1
