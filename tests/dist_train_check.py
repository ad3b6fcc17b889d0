"""Multi-rank check of UNetTrainer.training_step's bucketed, overlapped gradient all-reduce (needs >= 2 GPUs; run by hand):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_train_check.py

Every rank trains on its own data; after a (graph-replayed) step the flat gradient buffer of every rank must hold the SUM
over the ranks of the single-rank gradients -- computed here by a second, non-distributed trainer on each rank and one plain
all-reduce -- and the parameters of all ranks must stay identical."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import synth, train  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(1000 + rank)
N, S = 4, 32
x_start, cond = torch.randn(N, 8, S, S, generator=g), torch.randn(N, 8, S, S, generator=g)
feats, noise = torch.rand(N, 1, S, S, generator=g), torch.randn(N, 8, S, S, generator=g)
t = torch.randint(0, 1000, (N,), generator=g)
sd = synth.synth_unet_state(seed=0)
tr = train.UNetTrainer(sd, **synth.UNET_KWARGS, lr=1e-4, device=dev)
worst = 0.0
for step in range(3):                      # step 0 eager, step 1 captures, step 2 replays
    before = {k: v.clone() for k, v in tr.state_dict().items()}
    solo = train.UNetTrainer(before, **synth.UNET_KWARGS, lr=1e-4, device=dev)
    x = torch.cat([solo.scheduler.q_sample(x_start.to(dev), t.to(dev), noise.to(dev)), cond.to(dev), feats.to(dev)], dim=1)
    solo.forward_backward(x, t, noise.to(dev))
    want = solo.opt.grad.clone()
    dist.all_reduce(want)
    tr.training_step(x_start, cond, feats, t, noise)
    torch.cuda.synchronize()
    err = ((tr.opt.grad - want).norm() / want.norm()).item()
    worst = max(worst, err)
    # every rank applied the same update
    p = tr.opt.param.clone()
    ref = p.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(p, ref)
    if rank == 0:
        print(f"step {step}: reduced-gradient rel-L2 vs sum of single-rank gradients {err:.2e}, graphs {'on' if tr._graph['fb'] else 'off'}")
    assert err <= 1e-5, err          # fp32 atomics in wgrad: the two trainers differ in the last bits
    assert same, "parameters diverged between ranks"
if rank == 0:
    print(f"OK: {world} ranks, worst {worst:.2e}")
dist.destroy_process_group()
