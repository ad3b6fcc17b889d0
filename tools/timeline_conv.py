"""Per-CTA timeline of ONE planned convolution (debug build of the engine with -DB2D_TIMELINE, kept next to libb2d.so as
tools/probe/libb2d_timeline.so): where the kernel's duration goes -- launch ramp, set-up, first operands, K loop, accumulator
drain, split-K park / ticket / fix-up, tail.
usage: python tools/timeline_conv.py build                                  (here, no GPU needed)
       python tools/timeline_conv.py kind N cin cout H [block_n] [tune_flags] [ksplit] [ablation]     (on the GPU box)
ablation (generic staging): 1 = issue no MMAs, 2 = load no A tiles, 4 = load no B tiles (results are garbage, timing is the point)
kinds as in tools/profile_conv.py (2d / 3d / 1x1 / convT)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TL_LIB = os.path.join(ROOT, "tools", "probe", "libb2d_timeline.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    from diffusion_model_project_b200 import build as b
    print(b.build(force=True, extra_flags=("-DB2D_TIMELINE",), lib_path=TL_LIB, obj_dir=os.path.join("build", "timeline")))
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from diffusion_model_project_b200 import _lib  # noqa: E402

_lib.LIB_PATH = TL_LIB  # before the first call: this process runs the instrumented engine
from diffusion_model_project_b200 import engine  # noqa: E402
from diffusion_model_project_b200.engine import ConvPlan, new_act  # noqa: E402

a = sys.argv[1:]
kind, N, cin, cout, H = a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4])
bn = int(a[5]) if len(a) > 5 else 0
flags = int(a[6]) if len(a) > 6 else 0
ksplit = int(a[7]) if len(a) > 7 else 0
mode = int(a[8]) if len(a) > 8 else 0   # ablation: 1 = no MMAs, 2 = no A loads, 4 = no B loads (generic staging)
D = 11 if kind == "3d" else 1
dev = "cuda"
g = torch.Generator().manual_seed(0)
x = new_act(N, D, H, H, cin, dev, f16=True)
x.hi.view(torch.float16).copy_(torch.randn(N, D, H, H, cin, generator=g).to(torch.float16))
groups, use_stats = 1, 1
if kind == "3d":
    pw = engine.pack_conv3d(torch.randn(cout, cin, 3, 3, 3, generator=g) * (27 * cin) ** -0.5, torch.zeros(cout), dev, f16=True)
    groups = 32
elif kind == "convT":
    pw = engine.pack_convT2x2(torch.randn(cin, cout, 2, 2, generator=g) * cin ** -0.5, torch.zeros(cout), dev, f16=True)
elif kind == "1x1":
    pw = engine.pack_linear(torch.randn(cout, cin, generator=g) * cin ** -0.5, torch.zeros(cout), dev, f16=True)
    use_stats = 0
else:
    pw = engine.pack_conv2d(torch.randn(cout, cin, 3, 3, generator=g) * (9 * cin) ** -0.5, [cin], None, dev, f16=True)
up = 2 if kind == "convT" else 1
out = new_act(N, D, H * up, H * up, cout, dev, f16=True)
st = torch.zeros(N, groups, 2, dtype=torch.float64, device=dev)
plan = ConvPlan([x], pw, out, cout=cout, nphase=4 if kind == "convT" else 1, stats=st if use_stats else None,
                stats_cpg=cout // groups if use_stats else 0, block_n=bn, tune_flags=flags, tune_ksplit=ksplit)
s = torch.cuda.current_stream().cuda_stream
lib = _lib.lib()
lib.b2d_debug_timeline.restype = C.c_int
lib.b2d_debug_timeline.argtypes = [C.c_void_p, C.c_int]
assert lib.b2d_debug_timeline(None, 16 + mode) == 0
for _ in range(3):
    plan.run(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    plan.run(s)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100.0
assert lib.b2d_debug_timeline(None, 1) == 0
plan.run(s)
torch.cuda.synchronize()
SLOTS = 16
buf = np.zeros(296 * SLOTS * 2, dtype=np.uint64)
assert lib.b2d_debug_timeline(buf.ctypes.data, 0) == 0
tl = buf.reshape(296, SLOTS, 2).astype(np.int64)
live = tl[:, 0, 0] > 0
gt, ck = tl[live, :, 0], tl[live, :, 1]
n = gt.shape[0]
t0 = gt[:, 0].min()
names = {0: "entry", 1: "set-up done", 2: "first operands landed", 3: "last MMA issued", 9: "epilogue warp ready", 4: "accumulator complete",
         5: "partial parked", 6: "ticket taken", 7: "epilogue / fix-up done", 8: "all loads issued", 10: "all roles done"}
if mode:
    print(f"ABLATION mode {mode}: " + ", ".join(n_ for b_, n_ in ((1, "no MMAs"), (2, "no A loads"), (4, "no B loads")) if mode & b_))
print(f"conv{kind} {cin}->{cout} N={N} {D}x{H}x{H} {plan.info2()} : {us:.1f} us/launch (10 back-to-back launches, instrumented build), {n} CTAs")
print("globaltimer (ns since the first CTA's entry): min / median / max over CTAs that passed the point")
for slot in (0, 1, 2, 8, 3, 9, 4, 5, 6, 7, 10):
    ok = gt[:, slot] > 0
    if not ok.any():
        continue
    v = gt[ok, slot] - t0
    print(f"  {names[slot]:26s} n={ok.sum():3d}  {v.min():7d} {int(np.median(v)):7d} {v.max():7d}")
print("per-CTA intervals in SM clocks (clock64): median / max")


def iv(a_, b_, label):
    ok = (ck[:, a_] > 0) & (ck[:, b_] > 0)
    if ok.any():
        d = ck[ok, b_] - ck[ok, a_]
        print(f"  {label:44s} n={ok.sum():3d}  {int(np.median(d)):7d} {d.max():7d}")


iv(0, 1, "set-up (barriers, TMEM alloc, descriptor prefetch)")
iv(1, 2, "set-up done -> first operands landed")
iv(2, 3, "first operands -> last MMA issued (K loop)")
iv(3, 4, "last MMA issued -> accumulator complete")
iv(4, 5, "accumulator complete -> partial parked")
iv(5, 6, "parked -> ticket taken")
iv(6, 7, "ticket -> fix-up + epilogue done (last piece)")
iv(4, 7, "accumulator complete -> epilogue done")
iv(0, 10, "entry -> all roles done")
it = np.zeros(3 * 128, dtype=np.int64)
assert lib.b2d_debug_timeline(it.ctypes.data, -1) == 0
it = it.reshape(3, 128)
if it[0, 0] > 0 and it[1, 0] > 0:
    base = it[0, 0]
    print("CTA 0, K iteration: producer holds free stages | MMA warp sees full stages | commits issued   (SM clocks since the producer's first)")
    for i in range(min(28, int((it[0] > 0).sum()))):
        print(f"  {i:3d}  {it[0, i] - base:7d}  {it[1, i] - base:7d}  {it[2, i] - base:7d}")
print("note: K-loop clocks / k-groups of the unit = clocks per 64-wide K step; a 128 x BN x 64 step is 4 MMAs of BN/2 clocks")
