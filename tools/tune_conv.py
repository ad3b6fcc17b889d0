"""Sweep (BLOCK_N, K splits) for every distinct conv shape of one UNet step and print the measured time next to what the
plan's cost model picks -- the data for recalibrating the model in csrc/conv_plan.cu (b2d_conv_plan_create).
usage (on a GPU box): python tools/tune_conv.py [N] [reps]          N = slice-images (default 88 = 8 samples)
Forces the variants through b2d_conv_desc.block_n / tune_ksplit; infeasible combinations are skipped."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import engine  # noqa: E402
from diffusion_model_project_b200._lib import B2DError  # noqa: E402
from diffusion_model_project_b200.engine import ConvPlan, new_act  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 88
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = "cuda"
# (kind, cin segments, cout, H): the UNet's layer classes at a 64x64 latent
F_ = (64, 128, 256, 512, 1024)
shapes = []
for lvl, c in enumerate(F_):
    h = 64 >> lvl
    shapes += [("3x3", [c // 2 if lvl else 64], c, h), ("3x3", [c], c, h), ("3x3", [c, c], c, h), ("convT", [2 * c], c, h // 2)]
    if lvl >= 2:
        shapes += [("1x1", [c], 3 * c, h), ("1x1", [c], c, h)]
shapes += [("3x3", [1024], 2048, 2), ("3x3", [2048], 2048, 2)]


def timed(plan):
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.run(s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        plan.run(s)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


for kind, cins, cout, H in shapes:
    g = torch.Generator().manual_seed(0)
    xs = []
    for ci in cins:
        x = new_act(N, 1, H, H, ci, dev, f16=True)
        x.hi.view(torch.float16).copy_(torch.randn(N, 1, H, H, ci, generator=g).to(torch.float16))
        xs.append(x)
    cin = sum(cins)
    nphase, up = 1, 1
    if kind == "3x3":
        pw = engine.pack_conv2d(torch.randn(cout, cin, 3, 3, generator=g) * (9 * cin) ** -0.5, cins, None, dev, f16=True)
    elif kind == "convT":
        pw = engine.pack_convT2x2(torch.randn(cin, cout, 2, 2, generator=g) * cin ** -0.5, torch.zeros(cout), dev, f16=True)
        nphase, up = 4, 2
    else:
        pw = engine.pack_linear(torch.randn(cout, cin, generator=g) * cin ** -0.5, torch.zeros(cout), dev, f16=True)
    out = new_act(N, 1, H * up, H * up, cout, dev, f16=True)
    st = torch.zeros(N, 1, 2, dtype=torch.float64, device=dev) if kind != "1x1" else None
    kw = dict(cout=cout, nphase=nphase, stats=st, stats_cpg=cout if st is not None else 0)
    auto = ConvPlan(xs, pw, out, tune_flags=int(sys.argv[3]) if len(sys.argv) > 3 else 0, **kw)
    t_auto, info = timed(auto), auto.info2()
    rows = []
    for bn in (64, 128, 256):
        for ks in (1, 2, 3, 4, 6, 8):
            try:
                p = ConvPlan(xs, pw, out, block_n=bn, tune_ksplit=ks, tune_flags=32, **kw)
            except (B2DError, ValueError, RuntimeError):
                continue
            i2 = p.info2()
            if i2["block_n"] != bn or i2["ksplit"] != ks:
                continue
            rows.append((timed(p), bn, ks, i2["units"]))
    for bn in (64, 128, 256):  # stream-K at a forced tile width
        try:
            p = ConvPlan(xs, pw, out, block_n=bn, tune_flags=16, **kw)
        except (B2DError, ValueError, RuntimeError):
            continue
        i2 = p.info2()
        if i2["block_n"] == bn and i2["ksplit"] == -1:
            rows.append((timed(p), bn, -1, i2["units"]))
    for bn in (128, 256):  # tcgen05 cta_group::2 CTA pairs, with the planner's own K split and with forced ones
        for ks in (0, 1, 2, 3, 4, 6):
            try:
                p = ConvPlan(xs, pw, out, block_n=bn, tune_ksplit=ks, tune_flags=64 | 32, **kw)
            except (B2DError, ValueError, RuntimeError):
                continue
            i2 = p.info2()
            if i2["block_n"] == bn and i2["engine"] == 3 and (ks == 0 or i2["ksplit"] == ks):
                rows.append((timed(p), bn, 100 + i2["ksplit"], i2["units"]))
    rows = sorted(set(rows))
    best = rows[0] if rows else (float("nan"), 0, 0, 0)
    print(f"{kind:5s} {'+'.join(map(str, cins)):>9s}->{cout:<4d} @{H:<2d}  auto bn{info['block_n']} ks{info['ksplit']} halo{info['halo']} "
          f"{t_auto:6.1f} us{' PAIR' if info['engine'] == 3 else ''} | best bn{best[1]} ks{best[2]} ({best[3]} units) {best[0]:6.1f} us | "
          + "  ".join(f"bn{b}/{'sk' if k < 0 else ('pair-ks' + str(k - 100)) if k >= 100 else 'ks' + str(k)}:{t:.1f}" for t, b, k, _ in rows[:8]), flush=True)
