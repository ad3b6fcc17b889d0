"""Smallest program that launches one planned convolution of the sampling path, for quick timing of engine /
tile variants and as the target of `ncu --set full -k regex:conv_`.
usage: python tools/profile_conv.py kind N cin cout H [block_n] [tune_flags] [reps]
  kind = 3d  : Conv3d 3x3x3 on [N][11][H][H][cin]   (VAE; default 8 128 128 256 = D3D res3 conv)
  kind = 2d  : Conv2d 3x3 on [N][1][H][H][cin]      (UNet; e.g. 88 64 64 64 = encoder.0 block2)
  kind = 1x1 : Linear on [N][1][H][H][cin]          (UNet attention projections; e.g. 88 256 768 16 = enc2 in_proj)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import engine  # noqa: E402
from diffusion_model_project_b200.engine import ConvPlan, new_act  # noqa: E402

a = sys.argv[1:]
kind = a[0] if len(a) > 0 else "3d"
N = int(a[1]) if len(a) > 1 else 8
cin = int(a[2]) if len(a) > 2 else 128
cout = int(a[3]) if len(a) > 3 else 128
H = int(a[4]) if len(a) > 4 else 256
bn = int(a[5]) if len(a) > 5 else 0
flags = int(a[6]) if len(a) > 6 else 0  # b2d_conv_desc.tune_flags (B2D_TUNE_*)
reps = int(a[7]) if len(a) > 7 else 5
use_stats = int(a[8]) if len(a) > 8 else 1
D = 11 if kind == "3d" else 1
dev = "cuda"
g = torch.Generator().manual_seed(0)
x = new_act(N, D, H, H, cin, dev, f16=True)   # IEEE fp16 operands: the default 16-bit mode
x.hi.view(torch.float16).copy_(torch.randn(N, D, H, H, cin, generator=g).to(torch.float16))
if kind == "3d":
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) * (27 * cin) ** -0.5
    pw = engine.pack_conv3d(w, torch.zeros(cout), dev, f16=True)
    groups = 32
elif kind == "convT":  # ConvTranspose2d k2 s2 (UNet Up): 4 phase GEMMs, bias, GroupNorm(1, C) sums
    w = torch.randn(cin, cout, 2, 2, generator=g) * cin ** -0.5
    pw = engine.pack_convT2x2(w, torch.zeros(cout), dev, f16=True)
    groups = 1
elif kind == "1x1":  # attention projections (UNet in_proj / out_proj): bias, no GroupNorm sums
    w = torch.randn(cout, cin, generator=g) * cin ** -0.5
    pw = engine.pack_linear(w, torch.zeros(cout), dev, f16=True)
    groups = 1
    use_stats = 0
else:
    w = torch.randn(cout, cin, 3, 3, generator=g) * (9 * cin) ** -0.5
    pw = engine.pack_conv2d(w, [cin], None, dev, f16=True)
    groups = 1
up = 2 if kind == "convT" else 1
out = new_act(N, D, H * up, H * up, cout, dev, f16=True)
st = torch.zeros(N, groups, 2, dtype=torch.float64, device=dev)
plan = ConvPlan([x], pw, out, cout=cout, nphase=4 if kind == "convT" else 1, stats=st if use_stats else None,
                stats_cpg=cout // groups if use_stats else 0, block_n=bn, tune_flags=flags)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    plan.run(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    plan.run(s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"stats={use_stats} conv{kind} {cin}->{cout} N={N} {D}x{H}x{H} {plan.info2()} : {ms * 1e3:.1f} us/launch, {plan.flops / ms / 1e9:.1f} TFLOP/s")
