"""Oracle restatement of the dual-branch VAE's inference branches (E2D encoder, D3D decoder).

Test infrastructure (see oracle/__init__.py).
  Encoder.forward  VAE_model/src/vae/encoder.py:83-145
  Decoder.forward  VAE_model/src/vae/decoder.py:79-151
  ResidualBlock    VAE_model/src/vae/blocks.py:136-186
  DualBranchVAE.encode_2d_deterministic / decode_3d   VAE_model/src/dual_vae/model.py:211-233
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _res(sd, p, x):
    """blocks.py:173-186: x + conv2(SiLU(GN32(conv1(SiLU(GN32(x)))))) (+1x1x1 skip conv)."""
    h = F.silu(F.group_norm(x, 32, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], eps=1e-5))
    h = F.conv3d(h, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, 32, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], eps=1e-5))
    h = F.conv3d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.residual_layer.weight" in sd:
        x = F.conv3d(x, sd[f"{p}.residual_layer.weight"], sd[f"{p}.residual_layer.bias"])
    return h + x


def _down(sd, p, x):
    """encoder.py:76-81 + :45: F.pad(0,1,0,1,1,1) then Conv3d k3 stride (1,2,2) pad 0."""
    x = F.pad(x, (0, 1, 0, 1, 1, 1))
    return F.conv3d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], stride=(1, 2, 2))


def encoder_forward(sd, x, prefix="encoder_2d.", taps=None):
    """encoder.py:83-145 (conditional=False).  Returns (mu, logvar)."""
    p = prefix
    x = F.conv3d(x, sd[f"{p}conv_in.weight"], sd[f"{p}conv_in.bias"], padding=1)
    x = _res(sd, f"{p}res1_1", x)
    x = _res(sd, f"{p}res1_2", x)
    x = _down(sd, f"{p}down1", x)
    if taps is not None:
        taps["down1"] = x
    x = _res(sd, f"{p}res2_1", x)
    x = _res(sd, f"{p}res2_2", x)
    x = _down(sd, f"{p}down2", x)
    x = _res(sd, f"{p}res3_1", x)
    x = _res(sd, f"{p}res3_2", x)
    x = F.silu(F.group_norm(x, 32, sd[f"{p}norm_out.weight"], sd[f"{p}norm_out.bias"], eps=1e-5))
    x = F.conv3d(x, sd[f"{p}conv_out.weight"], sd[f"{p}conv_out.bias"], padding=1)
    mu, logvar = torch.chunk(x, 2, dim=1)
    return mu, logvar


def decoder_forward(sd, z, prefix="decoder_3d.", taps=None):
    """decoder.py:79-151 (conditional=False)."""
    p = prefix
    x = F.conv3d(z, sd[f"{p}conv_in.weight"], sd[f"{p}conv_in.bias"], padding=1)
    x = _res(sd, f"{p}res1_1", x)
    x = _res(sd, f"{p}res1_2", x)
    x = F.interpolate(x, scale_factor=(1, 2, 2))  # nn.Upsample default mode='nearest'
    x = F.conv3d(x, sd[f"{p}conv_up1.weight"], sd[f"{p}conv_up1.bias"], padding=1)
    if taps is not None:
        taps["up1"] = x
    x = _res(sd, f"{p}res2_1", x)
    x = _res(sd, f"{p}res2_2", x)
    x = F.interpolate(x, scale_factor=(1, 2, 2))
    x = F.conv3d(x, sd[f"{p}conv_up2.weight"], sd[f"{p}conv_up2.bias"], padding=1)
    x = _res(sd, f"{p}res3_1", x)
    x = _res(sd, f"{p}res3_2", x)
    x = F.silu(F.group_norm(x, 32, sd[f"{p}norm_out.weight"], sd[f"{p}norm_out.bias"], eps=1e-5))
    return F.conv3d(x, sd[f"{p}conv_out.weight"], sd[f"{p}conv_out.bias"], padding=1)


def encode_2d_deterministic(sd, x):
    """dual_vae/model.py:225-233: z = mu; logvar clamped to [-10, 10]."""
    mu, logvar = encoder_forward(sd, x, "encoder_2d.")
    return mu, (mu, torch.clamp(logvar, -10.0, 10.0))


def decode_3d(sd, z):
    """dual_vae/model.py:211-223."""
    return decoder_forward(sd, z, "decoder_3d.")
