"""Smallest program that launches the dominant kernel of the sampling path in its dominant shape:
conv_igemm_kernel<128,...> on D3D's 128->128 3x3x3 conv at 11x256x256 (B samples).  Used under
`ncu --set full -k regex:conv_igemm` and for quick timing of tile/stage variants.
usage: python tools/profile_conv.py [B] [cin] [cout] [H] [block_n] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import engine  # noqa: E402
from diffusion_model_project_b200.engine import ConvPlan, new_act  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cout = int(sys.argv[3]) if len(sys.argv) > 3 else 128
H = int(sys.argv[4]) if len(sys.argv) > 4 else 256
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
D = 11
dev = "cuda"
g = torch.Generator().manual_seed(0)
x = new_act(B, D, H, H, cin, dev)
x.hi.copy_(torch.randn(B, D, H, H, cin, generator=g).to(torch.bfloat16))
w = torch.randn(cout, cin, 3, 3, 3, generator=g) * (27 * cin) ** -0.5
pw = engine.pack_conv3d(w, torch.zeros(cout), dev)
out = new_act(B, D, H, H, cout, dev)
st = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
plan = ConvPlan([x], pw, out, cout=cout, stats=st, stats_cpg=cout // 32, block_n=bn)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    plan.run(s)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    plan.run(s)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"conv3d {cin}->{cout} B={B} {D}x{H}x{H} {plan.info()} : {ms:.3f} ms/launch, {plan.flops / ms / 1e9:.1f} TFLOP/s")
