"""Architecture spec (state-dict keys and shapes) and deterministic synthetic weights.

There is no network in the build/bench environment, so the shipped checkpoints
(Zenodo) are unavailable.  Throughput and parity runs therefore use random-init
weights "of the named architecture" (BASELINE.json).  This module enumerates the
reference's state-dict keys and shapes *without importing the reference* and
fills them from a seeded generator, so the reference (in the build container),
the CPU oracle and the CUDA path can all be loaded with bit-identical weights.

Key lists mirror:
  UNet                 Diffusion_model/src/unet/models.py:29-188, blocks.py:6-235
  Encoder / Decoder    VAE_model/src/vae/encoder.py:9-81, decoder.py:10-77, blocks.py:136-186
  DualBranchVAE        VAE_model/src/dual_vae/model.py:32-105
`tests/golden/reference_keys.json` (written by tests/golden/make_golden.py from the
real reference modules) pins this list; tests/test_synth.py checks it.

Zero-initialised modules (`final_conv`, every attention `proj_out`;
unet/models.py:120-128, unet/blocks.py:201-207) are re-randomised with
std 0.02, otherwise a random-init UNet predicts exactly 0.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch

UNET_KWARGS = dict(
    in_channels=17,
    out_channels=8,
    features=[64, 128, 256, 512, 1024],
    kernel_size=3,
    padding_mode="zeros",
    activation="silu",
    final_activation=None,
    attention="3..2",
    dropout=0.0,
    time_embedding_dim=64,
)

NORM_FACTORS = [0.01, 0.005, 0.002]


def attention_heads(expr: str, levels: int) -> List[int | None]:
    """`start.end.heads` expression -> heads per level (unet/models.py:316-366)."""
    out: List[int | None] = [None] * levels
    expr = (expr or "").strip()
    if not expr:
        return out
    try:
        a, b, h = expr.split(".")
        if not b.strip():
            b = str(levels)
        a, b, h = int(a) - 1, int(b) - 1, int(h)
        for i in range(a, b + 1):
            out[i] = h
    except Exception as exc:  # same error type/message class as the reference
        raise ValueError("Check validity of expression string.") from exc
    return out


def unet_spec(
    in_channels: int = 17,
    out_channels: int = 8,
    features: Sequence[int] = (64, 128, 256, 512, 1024),
    kernel_size: int = 3,
    attention: str = "3..2",
    time_embedding_dim: int | None = 64,
    **_unused,
) -> "OrderedDict[str, Tuple[int, ...]]":
    """Ordered {key: shape} of UNet.state_dict() (registration order of the reference)."""
    k = kernel_size
    feats = list(features)
    heads = attention_heads(attention, len(feats))
    tdim = None if time_embedding_dim is None else 4 * time_embedding_dim
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def block(prefix: str, cin: int, cout: int):
        spec[f"{prefix}.conv.weight"] = (cout, cin, k, k)
        spec[f"{prefix}.norm.weight"] = (cout,)
        spec[f"{prefix}.norm.bias"] = (cout,)

    def double(prefix: str, cin: int, cmid: int, cout: int):
        block(f"{prefix}.block1", cin, cmid)
        block(f"{prefix}.block2", cmid, cout)
        if tdim is not None:
            spec[f"{prefix}.time_mlp.1.weight"] = (cmid, tdim)
            spec[f"{prefix}.time_mlp.1.bias"] = (cmid,)

    def attn(prefix: str, c: int):
        spec[f"{prefix}.norm.weight"] = (c,)
        spec[f"{prefix}.norm.bias"] = (c,)
        spec[f"{prefix}.mha.in_proj_weight"] = (3 * c, c)
        spec[f"{prefix}.mha.in_proj_bias"] = (3 * c,)
        spec[f"{prefix}.mha.out_proj.weight"] = (c, c)
        spec[f"{prefix}.mha.out_proj.bias"] = (c,)
        spec[f"{prefix}.proj_out.weight"] = (c, c, 1)
        spec[f"{prefix}.proj_out.bias"] = (c,)

    if tdim is not None:
        spec["time_mlp.0.weight"] = (tdim, time_embedding_dim)
        spec["time_mlp.0.bias"] = (tdim,)
        spec["time_mlp.2.weight"] = (tdim, tdim)
        spec["time_mlp.2.bias"] = (tdim,)
    cin = in_channels
    for lvl, c in enumerate(feats):
        double(f"encoder.{lvl}.0", cin, c, c)
        if heads[lvl] is not None:
            attn(f"encoder.{lvl}.1", c)
        spec[f"encoder.{lvl}.2.norm.weight"] = (c,)
        spec[f"encoder.{lvl}.2.norm.bias"] = (c,)
        cin = c
    double("bottleneck", feats[-1], 2 * feats[-1], 2 * feats[-1])
    rheads = list(reversed(heads))
    for lvl, c in enumerate(reversed(feats)):
        spec[f"decoder.{lvl}.0.conv.weight"] = (2 * c, c, 2, 2)
        spec[f"decoder.{lvl}.0.conv.bias"] = (c,)
        spec[f"decoder.{lvl}.0.norm.weight"] = (c,)
        spec[f"decoder.{lvl}.0.norm.bias"] = (c,)
        double(f"decoder.{lvl}.1", 2 * c, c, c)
        if rheads[lvl] is not None:
            attn(f"decoder.{lvl}.2", c)
    spec["final_conv.weight"] = (out_channels, feats[0], k, k)
    spec["final_conv.bias"] = (out_channels,)
    return spec


def _res(spec, prefix: str, cin: int, cout: int):
    spec[f"{prefix}.norm1.weight"] = (cin,)
    spec[f"{prefix}.norm1.bias"] = (cin,)
    spec[f"{prefix}.conv1.weight"] = (cout, cin, 3, 3, 3)
    spec[f"{prefix}.conv1.bias"] = (cout,)
    spec[f"{prefix}.norm2.weight"] = (cout,)
    spec[f"{prefix}.norm2.bias"] = (cout,)
    spec[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3, 3)
    spec[f"{prefix}.conv2.bias"] = (cout,)
    if cin != cout:
        spec[f"{prefix}.residual_layer.weight"] = (cout, cin, 1, 1, 1)
        spec[f"{prefix}.residual_layer.bias"] = (cout,)


def encoder_spec(in_channels: int = 3, latent_channels: int = 8, prefix: str = ""):
    """Ordered {key: shape} of vae.encoder.Encoder (non-conditional)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    p = prefix
    s[f"{p}conv_in.weight"] = (128, in_channels, 3, 3, 3)
    s[f"{p}conv_in.bias"] = (128,)
    _res(s, f"{p}res1_1", 128, 128)
    _res(s, f"{p}res1_2", 128, 128)
    s[f"{p}down1.weight"] = (128, 128, 3, 3, 3)
    s[f"{p}down1.bias"] = (128,)
    _res(s, f"{p}res2_1", 128, 256)
    _res(s, f"{p}res2_2", 256, 256)
    s[f"{p}down2.weight"] = (256, 256, 3, 3, 3)
    s[f"{p}down2.bias"] = (256,)
    _res(s, f"{p}res3_1", 256, 512)
    _res(s, f"{p}res3_2", 512, 512)
    s[f"{p}norm_out.weight"] = (512,)
    s[f"{p}norm_out.bias"] = (512,)
    s[f"{p}conv_out.weight"] = (2 * latent_channels, 512, 3, 3, 3)
    s[f"{p}conv_out.bias"] = (2 * latent_channels,)
    return s


def decoder_spec(latent_channels: int = 8, out_channels: int = 3, prefix: str = ""):
    """Ordered {key: shape} of vae.decoder.Decoder (non-conditional)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    p = prefix
    s[f"{p}conv_in.weight"] = (512, latent_channels, 3, 3, 3)
    s[f"{p}conv_in.bias"] = (512,)
    _res(s, f"{p}res1_1", 512, 512)
    _res(s, f"{p}res1_2", 512, 512)
    s[f"{p}conv_up1.weight"] = (256, 512, 3, 3, 3)
    s[f"{p}conv_up1.bias"] = (256,)
    _res(s, f"{p}res2_1", 256, 256)
    _res(s, f"{p}res2_2", 256, 256)
    s[f"{p}conv_up2.weight"] = (128, 256, 3, 3, 3)
    s[f"{p}conv_up2.bias"] = (128,)
    _res(s, f"{p}res3_1", 128, 128)
    _res(s, f"{p}res3_2", 128, 128)
    s[f"{p}norm_out.weight"] = (128,)
    s[f"{p}norm_out.bias"] = (128,)
    s[f"{p}conv_out.weight"] = (out_channels, 128, 3, 3, 3)
    s[f"{p}conv_out.bias"] = (out_channels,)
    return s


def dual_vae_spec(in_channels: int = 3, latent_channels: int = 8, branches=("encoder_2d", "decoder_2d", "encoder_3d", "decoder_3d")):
    """Ordered {key: shape} of DualBranchVAE.state_dict() (dual_vae/model.py:69-104)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for b in branches:
        if b.startswith("encoder"):
            s.update(encoder_spec(in_channels, latent_channels, prefix=b + "."))
        else:
            s.update(decoder_spec(latent_channels, in_channels, prefix=b + "."))
    return s


def _fan_in(shape: Tuple[int, ...], key: str) -> int:
    if key.endswith("decoder_conv_transpose"):
        return shape[0] * shape[2] * shape[3]
    if len(shape) == 1:
        return shape[0]
    f = 1
    for d in shape[1:]:
        f *= d
    return f


def synth_state_dict(spec: Dict[str, Tuple[int, ...]], seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Fill `spec` deterministically.

    weights  ~ U(-b, b), b = sqrt(3 / fan_in) (unit-gain, keeps activations O(1) through
               the GroupNorm-free residual sums of the VAE), drawn per tensor from
               torch.Generator().manual_seed(seed * 100003 + index) so the values do not
               depend on which other tensors exist;
    norm     weight = 1 + 0.1 u, bias = 0.1 u  (so the affine is exercised);
    biases   0.1 u;
    zero-init modules of the reference (final_conv, *.proj_out) ~ N(0, 0.02^2).
    """
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for idx, (key, shape) in enumerate(spec.items()):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        leaf = key.rsplit(".", 1)[-1]
        is_norm = ".norm" in key or "norm_out" in key
        if key.startswith("final_conv.weight") or key.endswith("proj_out.weight"):
            t = torch.randn(shape, generator=g, dtype=torch.float32) * 0.02
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * (2 * torch.rand(shape, generator=g) - 1)
        elif leaf in ("bias", "in_proj_bias") or (is_norm and leaf == "bias"):
            t = 0.1 * (2 * torch.rand(shape, generator=g) - 1)
        else:
            if len(shape) == 4 and shape[2] == 2 and ".0.conv.weight" in key and key.startswith("decoder."):
                fan = shape[0]  # ConvTranspose2d k2 s2: each output pixel sees Cin taps
            else:
                fan = _fan_in(shape, key)
            b = math.sqrt(3.0 / fan)
            t = (2 * torch.rand(shape, generator=g, dtype=torch.float32) - 1) * b
        out[key] = t.to(dtype)
    return out


def synth_unet_state(seed: int = 0, **kwargs):
    kw = dict(UNET_KWARGS)
    kw.update(kwargs)
    return synth_state_dict(unet_spec(**kw), seed=seed)


def synth_vae_state(seed: int = 1, in_channels: int = 3, latent_channels: int = 8, branches=("encoder_2d", "decoder_3d")):
    """Only the two branches on the inference path by default (E2D, D3D)."""
    return synth_state_dict(dual_vae_spec(in_channels, latent_channels, branches), seed=seed)


def synth_inputs(batch: int, num_slices: int = 11, size: int = 256, seed: int = 2024):
    """Synthetic microstructure + 2D velocity."""
    g = torch.Generator().manual_seed(seed)
    img = (torch.rand(batch, num_slices, 1, size, size, generator=g) > 0.4).float()
    v2d = torch.randn(batch, num_slices, 3, size, size, generator=g) * 0.003 * img
    v2d[:, :, 2] = 0
    return img, v2d


def synth_noise(batch: int, num_slices: int = 11, latent_channels: int = 8, latent_size: int = 64, seed: int = 42):
    """Initial latent noise, one generator per sample (eval_testset_end2end.py:809-810)."""
    out = []
    for i in range(batch):
        g = torch.Generator().manual_seed(seed + i)
        out.append(torch.randn(num_slices, latent_channels, latent_size, latent_size, generator=g))
    return torch.cat(out, 0)
