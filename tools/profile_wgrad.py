"""One weight-gradient launch of the training step (train.conv_wgrad, channels-last dW) at a UNet layer's shape: the ncu target.
usage: python tools/profile_wgrad.py [N] [H] [cout] [cin]      (default 22 2 2048 2048 = bottleneck.block2)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import train  # noqa: E402
from diffusion_model_project_b200.engine import new_act  # noqa: E402

a = sys.argv[1:]
N, H, co, ci = (int(a[i]) if len(a) > i else v for i, v in enumerate((22, 2, 2048, 2048)))
dev = "cuda"
dy, x = new_act(N, 1, H, H, co, dev, split=True), new_act(N, 1, H, H, ci, dev, split=True)
for t in (dy, x):
    t.hi.copy_(torch.randn(t.hi.shape, device=dev).to(torch.bfloat16))
    t.lo.copy_((torch.randn(t.lo.shape, device=dev) * 1e-3).to(torch.bfloat16))
dw = torch.zeros(co, 3, 3, ci, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    train.conv_wgrad(dy, x, dw, co, ci, 0, s, channels_last=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    train.conv_wgrad(dy, x, dw, co, ci, 0, s, channels_last=True)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
flops = 2.0 * N * H * H * co * ci * 9 * 3       # three hi/lo products
print(f"conv_wgrad N={N} {H}x{H} {ci}->{co}: {us:.1f} us, dW {dw.numel() * 4 / 1e6:.1f} MB -> {dw.numel() * 4 / us / 1e3:.0f} GB/s of reductions, "
      f"{flops / us / 1e6:.1f} TFLOP/s of MMA work")
