"""Oracle restatement of the conditioned UNet (Diffusion_model/src/unet/models.py:131-188).

Test infrastructure (see oracle/__init__.py).  Functional fp32 CPU code driven by a flat
state dict with the reference's keys; no nn.Module from the reference is imported.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F


def sinusoidal_embedding(time: torch.Tensor, dim: int) -> torch.Tensor:
    """models.py:14-26."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, device=time.device) * -e)
    e = time[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def _block(sd, p, x):
    """blocks.py:6-47: conv3x3 (no bias, zero padding) -> GroupNorm(1,C) -> SiLU."""
    x = F.conv2d(x, sd[f"{p}.conv.weight"], None, padding=1)
    x = F.group_norm(x, 1, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-5)
    return F.silu(x)


def _double(sd, p, x, temb):
    """blocks.py:50-107."""
    x = _block(sd, f"{p}.block1", x)
    if temb is not None:
        tc = F.linear(F.silu(temb), sd[f"{p}.time_mlp.1.weight"], sd[f"{p}.time_mlp.1.bias"])
        x = x + tc[:, :, None, None]
    return _block(sd, f"{p}.block2", x)  # dropout p=0 -> identity


def _attention(sd, p, x, heads):
    """blocks.py:177-235: x + Conv1d_1x1(MHA(GN_1(x)^T)); nn.MultiheadAttention(batch_first) math."""
    b, c, h, w = x.shape
    xn = F.group_norm(x, 1, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-5)
    tok = xn.view(b, c, h * w).swapaxes(1, 2)  # (B, T, C)
    qkv = F.linear(tok, sd[f"{p}.mha.in_proj_weight"], sd[f"{p}.mha.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    d = c // heads
    T = h * w

    def split(t):
        return t.reshape(b, T, heads, d).transpose(1, 2)  # (B, heads, T, d)

    q, k, v = split(q), split(k), split(v)
    att = torch.softmax((q / math.sqrt(d)) @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(b, T, c)
    o = F.linear(o, sd[f"{p}.mha.out_proj.weight"], sd[f"{p}.mha.out_proj.bias"])
    o = o.swapaxes(2, 1)  # (B, C, T)
    hh = F.conv1d(o, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
    return x + hh.reshape(b, c, h, w)


def _down(sd, p, x):
    """blocks.py:146-174: maxpool2x2 -> GN(1,C) -> SiLU."""
    x = F.max_pool2d(x, 2, 2)
    x = F.group_norm(x, 1, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-5)
    return F.silu(x)


def _up(sd, p, x):
    """blocks.py:111-143: ConvTranspose2d k2 s2 (+bias) -> GN(1,C) -> SiLU."""
    x = F.conv_transpose2d(x, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], stride=2)
    x = F.group_norm(x, 1, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-5)
    return F.silu(x)


def time_embedding(sd, time: torch.Tensor, dim: int = 64) -> torch.Tensor:
    """models.py:141-142: sinusoid -> Linear -> SiLU -> Linear."""
    e = sinusoidal_embedding(time, dim)
    e = F.linear(e, sd["time_mlp.0.weight"], sd["time_mlp.0.bias"])
    return F.linear(F.silu(e), sd["time_mlp.2.weight"], sd["time_mlp.2.bias"])


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, time: torch.Tensor,
                 features: Sequence[int] = (64, 128, 256, 512, 1024), heads=(None, None, 2, 2, 2),
                 time_embedding_dim: int = 64, taps: dict | None = None) -> torch.Tensor:
    """models.py:131-188.  `taps` (optional dict) receives named intermediates for layer-wise tests."""
    temb = time_embedding(sd, time, time_embedding_dim)
    skips = []
    for lvl in range(len(features)):
        x = _double(sd, f"encoder.{lvl}.0", x, temb)
        if heads[lvl] is not None:
            x = _attention(sd, f"encoder.{lvl}.1", x, heads[lvl])
        if taps is not None:
            taps[f"enc{lvl}"] = x
        skips.append(x)
        x = _down(sd, f"encoder.{lvl}.2", x)
    x = _double(sd, "bottleneck", x, temb)
    if taps is not None:
        taps["bottleneck"] = x
    skips.reverse()
    rheads = list(reversed(heads))
    for lvl in range(len(features)):
        x = _up(sd, f"decoder.{lvl}.0", x)
        x = torch.cat((skips[lvl], x), dim=1)
        x = _double(sd, f"decoder.{lvl}.1", x, temb)
        if rheads[lvl] is not None:
            x = _attention(sd, f"decoder.{lvl}.2", x, rheads[lvl])
        if taps is not None:
            taps[f"dec{lvl}"] = x
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"], padding=1)  # final_activation None
