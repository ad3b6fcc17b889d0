"""Per-launch timing of one UNet step (N slice-images of 64x64 latent) with CUDA events, eager (no graph).
usage: python tools/profile_unet.py [N] [reps] [nograph|graph] [tune_flags]   -- also the target of the ncu launch-list pass.
tune_flags: _lib.TUNE_* bits for every conv plan (16 = stream-K wherever it applies, 32 = never)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import synth  # noqa: E402
from diffusion_model_project_b200.unet import B200UNet  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 88
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.set_grad_enabled(False)
m = B200UNet(**synth.UNET_KWARGS, device="cuda").load_state_dict(synth.synth_unet_state(seed=0))
m.conv_tune_flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m.fuse_block1_norm = len(sys.argv) > 5 and sys.argv[5] == "fuse1"   # block1's norm applied inside block2's conv
st = m.build_program(N, 64, 64)
st["x_in"].hi.view(torch.float16).copy_(torch.randn(N, 1, 64, 64, 64, device="cuda").to(torch.float16))
prog = st["program"]
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    prog.run(s)
torch.cuda.synchronize()
names = [n for n, _ in prog.steps]
acc = [0.0] * len(names)
for _ in range(reps):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    evs[0].record()
    for i, (_, fn) in enumerate(prog.steps):
        fn(s)
        evs[i + 1].record()
    torch.cuda.synchronize()
    for i in range(len(names)):
        acc[i] += evs[i].elapsed_time(evs[i + 1]) * 1e3 / reps
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    prog.run(s)
b.record()
torch.cuda.synchronize()
total = a.elapsed_time(b) * 1e3 / reps
print(f"UNet step N={N}: {total:.1f} us back-to-back, {sum(acc):.1f} us summed per-launch, {len(names)} launches, "
      f"{prog.flops / total / 1e6:.1f} TFLOP/s")
plans = {}
for k in st["keep"]:
    if hasattr(k, "info"):
        plans[id(k)] = k
for (n, fn), t in zip(prog.steps, acc):
    owner = getattr(fn, "__self__", None)
    extra = ""
    if owner is not None and hasattr(owner, "info2"):
        i2 = owner.info2()
        extra = (f"  {owner.flops / t / 1e6:7.1f} TFLOP/s  halo{i2['halo']} bn{i2['block_n']} ks{i2['ksplit']} units{i2['units']} ctas{i2['ctas']} "
                 f"kgroups{i2['kgroups']}")
    print(f"{t:9.1f} us  {n}{extra}")
kinds = {}
for n, t in zip(names, acc):
    k = n.rsplit(".", 1)[-1]
    kinds[k] = kinds.get(k, 0.0) + t
print({k: round(v, 1) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])})
if len(sys.argv) > 3 and sys.argv[3] == "nograph":  # under ncu: the eager launches are the ones to list
    sys.exit(0)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    prog.run(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
a.record()
for _ in range(reps * 4):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f"graph-replayed step: {a.elapsed_time(b) * 1e3 / (reps * 4):.1f} us ({len(names)} launches)")
