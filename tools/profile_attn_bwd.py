"""One attention-core backward (train.attention_bwd) at a UNet level's shape: the ncu target.
usage: python tools/profile_attn_bwd.py [N] [T] [C] [heads]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import train  # noqa: E402
from diffusion_model_project_b200.engine import new_act  # noqa: E402

a = sys.argv[1:]
N, T, C, heads = (int(a[i]) if len(a) > i else v for i, v in enumerate((22, 256, 256, 2)))
dev = "cuda"
mk = lambda c: new_act(N, 1, 1, T, c, dev, split=True)
qkv, out, dout, dqkv = mk(3 * C), mk(C), mk(C), mk(3 * C)
for t in (qkv, out, dout):
    t.hi.copy_(torch.randn(t.hi.shape, device=dev).to(torch.bfloat16))
    t.lo.copy_((torch.randn(t.lo.shape, device=dev) * 1e-3).to(torch.bfloat16))
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    train.attention_bwd(qkv, out, dout, dqkv, heads, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    train.attention_bwd(qkv, out, dout, dqkv, heads, s)
e1.record()
torch.cuda.synchronize()
print(f"attention_bwd N={N} T={T} C={C} heads={heads}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us (3 launches)")
