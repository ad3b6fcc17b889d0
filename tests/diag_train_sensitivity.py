"""Diagnostic (CPU, not a test): how sensitive are the UNet's parameter gradients to the 2^-17 storage rounding of the fp32-class
mode?  Quantises ONLY the inputs of the five max-pools to bf16 hi + lo (straight-through gradient) inside the oracle's fp32
autograd step at BASELINE configs[4]'s per-GPU size (22 slice-images of 8 x 64 x 64) and compares every gradient tensor.

    python tests/diag_train_sensitivity.py

Result (recorded in tests/test_gpu_train.py::test_training_step_at_baseline_size): 3 of 2.8 M pooling windows change their
argmax, and that alone moves encoder.1.2.norm.weight by 2.1e-3, encoder.2.0.block1.conv.weight by 1.7e-3 (rel-L2; median
tensor 1.7e-5) -- the same tensors, at the same magnitude, that carry the largest error of the GPU step (3.2e-3, 2.8e-3;
median 8.7e-5).  The fp32 oracle itself is within 4.5e-6 of an fp64 run of the same graph.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import synth  # noqa: E402
from oracle import train as otrain, unet as ounet  # noqa: E402


def case(N=22, S=64, seed=77):
    g = torch.Generator().manual_seed(seed)
    x_start, cond = torch.randn(N, 8, S, S, generator=g), torch.randn(N, 8, S, S, generator=g)
    feats, noise = torch.rand(N, 1, S, S, generator=g), torch.randn(N, 8, S, S, generator=g)
    t = torch.randint(0, 1000, (N,), generator=g)
    return x_start, cond, feats, t, noise


def q16(x):
    hi = x.to(torch.bfloat16).float()
    return hi + (x - hi).to(torch.bfloat16).float()


def main():
    sd = synth.synth_unet_state(seed=0)
    x_start, cond, feats, t, noise = case()
    _, g32, _ = otrain.training_loss_and_grads(sd, x_start, cond, feats, t, noise)
    flips = []

    def down_q(sd_, p, x):
        xq = x + (q16(x.detach()) - x.detach())
        with torch.no_grad():
            i0 = F.max_pool2d(x, 2, 2, return_indices=True)[1]
            i1 = F.max_pool2d(xq, 2, 2, return_indices=True)[1]
            flips.append((p, int((i0 != i1).sum()), i0.numel()))
        y = F.group_norm(F.max_pool2d(xq, 2, 2), 1, sd_[f"{p}.norm.weight"], sd_[f"{p}.norm.bias"], eps=1e-5)
        return F.silu(y)

    ounet._down = down_q
    _, gq, _ = otrain.training_loss_and_grads(sd, x_start, cond, feats, t, noise)
    errs = {n: ((gq[n] - g32[n]).norm() / g32[n].norm()).item() for n in g32}
    print("argmax flips per pool:", flips)
    print("largest:", [(k, f"{v:.1e}") for k, v in sorted(errs.items(), key=lambda kv: -kv[1])[:8]])
    print("median:", sorted(errs.values())[len(errs) // 2])


if __name__ == "__main__":
    main()
