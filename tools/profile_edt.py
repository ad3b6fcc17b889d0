"""Time the exact Euclidean distance transform (b2d_edt2d) of one conditioning pass: 88 slice-images of 256 x 256.
usage: python tools/profile_edt.py [n_img] [size]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 88
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
imgs = (torch.rand(n, S, S, device="cuda") > 0.4).float()
out = torch.empty(2 * n * S * S, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.call("b2d_edt2d", imgs.data_ptr(), out.data_ptr(), n, S, S, s, launches=2)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    _lib.call("b2d_edt2d", imgs.data_ptr(), out.data_ptr(), n, S, S, s, launches=2)
b.record()
torch.cuda.synchronize()
print(f"b2d_edt2d {n} x {S} x {S}: {a.elapsed_time(b) / 10 * 1e3:.1f} us (column pass + row pass)")
