"""gn_apply on the largest VAE tensor (8 x 11 x 256 x 256 x 128 bf16, in place): HBM roofline of the elementwise pass."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import engine  # noqa: E402
from diffusion_model_project_b200.engine import new_act  # noqa: E402

dev = "cuda"
B, D, H, C = 8, 11, 256, 128
x = new_act(B, D, H, H, C, dev, f16=True)
x.hi.view(torch.float16).copy_(torch.randn(B, D, H, H, C, device=dev, dtype=torch.float16))
n = D * H * H * (C // 32)
st = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
st[..., 0] = 0.1 * n
st[..., 1] = 1.01 * n
g = torch.ones(C, device=dev)
b = torch.zeros(C, device=dev)
s = torch.cuda.current_stream().cuda_stream
y = x.as_bf16()
for _ in range(3):
    engine.gn_apply(x, y, st, C // 32, g, b, True, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    engine.gn_apply(x, y, st, C // 32, g, b, True, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
nbytes = 2 * x.hi.numel() * 2
print(f"gn_apply {B}x{D}x{H}x{H}x{C}: {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s ({nbytes / 1e9:.2f} GB read+write)")
