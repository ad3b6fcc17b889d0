"""Chain (one cooperative persistent kernel for the whole UNet step) against the launch-per-layer program on the same
buffers.  usage: python tools/diag_chain.py [N] [h] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import engine, synth  # noqa: E402
from diffusion_model_project_b200.unet import B200UNet  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
h = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.set_grad_enabled(False)
m = B200UNet(**synth.UNET_KWARGS, device="cuda").load_state_dict(synth.synth_unet_state(seed=0))
st = m.build_program(N, h, h, fuse_small=False)
st["x_in"].hi.copy_(torch.randn(N, 1, h, h, 64, device="cuda").to(torch.bfloat16))
st["x_in"].hi[..., 17:] = 0
prog = st["program"]
s = torch.cuda.current_stream().cuda_stream
prog.run(s)
torch.cuda.synchronize()
ref = st["eps"].clone()
chain = engine.Chain(prog, "cuda")
print(f"chain: {chain.num_ops} ops", flush=True)
st["eps"].zero_()
chain.run(s)
torch.cuda.synchronize()
got = st["eps"].clone()
err = ((got - ref).abs().max() / ref.abs().max()).item()
print(f"N={N} {h}x{h}: chain vs program max-rel diff {err:.3e}  (equal: {torch.equal(got, ref)})", flush=True)


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    prog.run(torch.cuda.current_stream().cuda_stream)
print(f"program: {timeit(lambda: prog.run(s)):.1f} us eager, {timeit(g.replay):.1f} us graph-replayed; chain: {timeit(lambda: chain.run(s)):.1f} us")

# per-op device times inside the chain next to the eager per-launch event times of the program
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
chain.run(s)
torch.cuda.synchronize()
ct = chain.op_times_us()
kinds = {}
for n, a, c in zip(names, acc, ct):
    k = n.rsplit(".", 1)[-1]
    e = kinds.setdefault(k, [0.0, 0.0, 0])
    e[0] += a; e[1] += c; e[2] += 1
    print(f"{a:8.1f} us eager launch | {c:8.1f} us in chain   {n}")
print({k: (round(v[0], 1), round(v[1], 1), v[2]) for k, v in kinds.items()})
