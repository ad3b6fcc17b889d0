"""Helpers shared by the GPU parity tests: channels-last packing of test tensors and a plain
PyTorch fp32 reference (TF32 off) of the op under test."""
from __future__ import annotations

import torch

from diffusion_model_project_b200 import engine
from diffusion_model_project_b200.engine import Act, pad64


def no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()


def f16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).float()


def fmt_round(x: torch.Tensor, f16: bool) -> torch.Tensor:
    return f16_round(x) if f16 else bf16_round(x)


def to_act(x: torch.Tensor, split=False, f16=False) -> Act:
    """planar fp32 (N,C,D,H,W) -> channels-last Act [N,D,H,W,pad64(C)] on x.device (bf16, or IEEE fp16 if f16)."""
    N, C, D, H, W = x.shape
    cp = pad64(C)
    cl = torch.zeros(N, D, H, W, cp, dtype=torch.float32, device=x.device)
    cl[..., :C] = x.permute(0, 2, 3, 4, 1)
    if f16:
        assert not split
        return Act(cl.to(torch.float16).contiguous().view(torch.bfloat16), None, True)
    hi = cl.to(torch.bfloat16).contiguous()
    lo = (cl - hi.float()).to(torch.bfloat16).contiguous() if split else None
    return Act(hi, lo)


def from_act(a: Act, C: int) -> torch.Tensor:
    """Act -> planar fp32 (N,C,D,H,W)."""
    v = a.hi.view(torch.float16).float() if a.f16 else a.hi.float()
    if a.lo is not None:
        v = v + a.lo.float()
    return v[..., :C].permute(0, 4, 1, 2, 3).contiguous()


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (the per-step metric of BASELINE.json's north_star)."""
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def stats_ref(y: torch.Tensor, groups: int):
    """y planar (N,C,D,H,W) -> (N, groups, 2) [sum, sumsq] in float64."""
    N, C = y.shape[:2]
    g = y.double().reshape(N, groups, -1)
    return torch.stack([g.sum(-1), (g * g).sum(-1)], dim=-1)
