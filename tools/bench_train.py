"""BASELINE.json configs[4]: one UNet training step per GPU on 2 samples (22 slice-images of 8x64x64 latents), gradients
all-reduced over NCCL, Adam, operand refresh (train.UNetTrainer.training_step).  One JSON line from rank 0.
usage: python tools/bench_train.py [--slices 22] [--size 64] [--steps 10] [--warmup 3]
       python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py
parts_ms: the step's pieces each timed ALONE after the timed loop (forward_backward = the three phase graphs back to back;
allreduce_alone = one all-reduce of the whole flat gradient); allreduce_exposed = ms_per_step minus the other pieces, i.e.
what the bucketed all-reduce adds to the step with its first two buckets running under the backward."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import _lib, synth, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slices", type=int, default=22)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--precision", default="fp32x", choices=["fp32x", "bf16"])
ap.add_argument("--eager", action="store_true", help="no CUDA graphs: every launch issued from Python")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, S = args.slices, args.size
g = torch.Generator().manual_seed(100 + rank)
x_start, cond = torch.randn(N, 8, S, S, generator=g).to(dev), torch.randn(N, 8, S, S, generator=g).to(dev)
feats, noise = torch.rand(N, 1, S, S, generator=g).to(dev), torch.randn(N, 8, S, S, generator=g).to(dev)
t = torch.randint(0, 1000, (N,), generator=g).to(dev)
tr = train.UNetTrainer(synth.synth_unet_state(seed=0), **synth.UNET_KWARGS, lr=1e-4, device=dev, precision=args.precision)


def ev():
    return torch.cuda.Event(enable_timing=True)


def step():
    return tr.training_step(x_start, cond, feats, t, noise, use_graph=not args.eager)[0]


def timed(fn, reps):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for _ in range(max(args.warmup, 2)):   # the second step of a shape captures the graphs
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
a, b = ev(), ev()
a.record()
for _ in range(args.steps):
    loss = step()
b.record()
host = (time.perf_counter() - t0) * 1e3 / args.steps   # launch-side time: if it equals the step time the host is the limiter
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / args.steps
ms = a.elapsed_time(b) / args.steps
# the step's pieces, each timed alone (in the step the first two all-reduce buckets run under the backward)
gr = tr._graph
parts = {}
if gr is not None and gr.get("fb"):
    parts["forward_backward"] = timed(lambda: [g.replay() for g in gr["fb"]], args.steps)
    parts["operand_refresh"] = timed(gr["refresh"].replay, args.steps)
else:
    parts["forward_backward"] = timed(lambda: tr._q_sample_phases and [None for _ in tr._q_sample_phases(x_start, cond, feats, t, noise)], args.steps)
    parts["operand_refresh"] = timed(tr.refresh_operands, args.steps)
parts["adam"] = timed(lambda: tr.opt.step(grad_scale=0.0), 3)   # grad_scale 0 and lr untouched: moments decay, parameters move by ~0
parts["allreduce_alone"] = timed(lambda: dist.all_reduce(tr.opt.grad), 5) if world > 1 else 0.0
parts["allreduce_exposed"] = max(0.0, ms - parts["forward_backward"] - parts["operand_refresh"] - parts["adam"])
tm = torch.tensor([ms], device=dev)
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    nparam = sum(v.numel() for v in tr.state_dict().values())
    # algorithmic work: forward 8.713 GFLOP per slice-image at 64x64, backward = 2x forward
    flops = 3 * 8.713e9 * N * (S / 64.0) ** 2
    print(json.dumps({
        "metric": "UNet training steps/sec (forward + backward + gradient all-reduce + Adam)", "value": 1e3 / tm.item() * 1.0,
        "unit": "steps/s", "n_gpus": world, "ms_per_step": tm.item(), "wall_ms_per_step": wall, "host_issue_ms_per_step": host, "slices_per_gpu": N, "latent": S,
        "samples_per_s": world * (N / 11.0) * 1e3 / tm.item(), "dtype": "fp32x (bf16 hi + lo operands, fp32 accumulate)" if args.precision == "fp32x" else "bf16 (fp32 accumulate)",
        "parts_ms": parts,
        "algorithmic_tflops": flops / (parts["forward_backward"] * 1e-3) / 1e12, "parameters": nparam, "gradient_bytes": 4 * tr.opt.numel,
        "allreduce_busbw_GBps": (2 * (world - 1) / world * 4 * tr.opt.numel / (parts["allreduce_alone"] * 1e-3) / 1e9) if world > 1 else None,
        "allreduce": "three buckets in backward order; the first two (decoder + final_conv, bottleneck) on a side stream under the backward",
        "gpu_launches_per_step": tr._fb_launches + 1 + sum(len(o.parts) for l in tr._layers.values() for o in (l.operands() if hasattr(l, "operands") else l)), "cuda_graphs": not args.eager, "loss": loss.item(), "data": "synthetic"}))
if world > 1:
    dist.destroy_process_group()
