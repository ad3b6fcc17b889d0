"""Oracle restatement of LatentDiffusionPredictor.predict / predict_ddim
(Diffusion_model/src/predictor.py:754-1023) and its glue (pre_process :1025-1040,
apply_distance_transform :1096-1116, MaxNormalizer normalizer.py:46-58).

Test infrastructure (see oracle/__init__.py).  The reference's wasted shape-probe E2D pass
(predictor.py:765-774) is not replayed: it has no effect on the result.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage

from . import unet as ounet
from . import vae as ovae
from .scheduler import OracleScheduler, ddim_timesteps


def distance_transform(imgs: torch.Tensor) -> torch.Tensor:
    """predictor.py:1096-1116: SciPy exact EDT per (n,1,H,W) image, float64 -> float32."""
    arr = imgs.cpu().numpy()
    out = [ndimage.distance_transform_edt(im[0])[None, None] for im in arr]
    return torch.from_numpy(np.concatenate(out)).float()


def conditioning(vae_sd, img, velocity_2d, norm_factors, use_edt=True):
    """predictor.py:927-962 (= :777-812).  Returns (v2d_latent (N,8,h,w), feats (N,1,h,w))."""
    B, S = img.shape[0], velocity_2d.shape[1]
    H, W = img.shape[3], img.shape[4]
    s = torch.tensor(norm_factors, dtype=torch.float32).view(1, 1, -1, 1, 1)
    v = (velocity_2d / s).permute(0, 2, 1, 3, 4)  # (B,3,S,H,W), normalizer.py:46-51
    z, _ = ovae.encode_2d_deterministic(vae_sd, v)  # (B,8,S,h,w)
    lc, ld, lh, lw = z.shape[1:]
    v_lat = z.permute(0, 2, 1, 3, 4).reshape(B * ld, lc, lh, lw)
    img_flat = img.reshape(B * S, 1, H, W)
    feats = distance_transform(img_flat) if use_edt else img_flat
    feats = feats / 1.0  # normalizer['input'] scale 1 (predictor.py:337-338)
    feats = F.interpolate(feats, size=(lh, lw), mode="bilinear", align_corners=False)
    f3 = feats.reshape(B, S, 1, lh, lw).permute(0, 2, 1, 3, 4)
    f3 = F.interpolate(f3, size=(ld, lh, lw), mode="trilinear", align_corners=False).permute(0, 2, 1, 3, 4)
    return v_lat, f3.reshape(B * ld, 1, lh, lw)


def encode_target(vae_sd, velocity_3d, norm_factors):
    """predictor.py:1042-1085: (B,S,3,H,W) -> permute, MaxNormalizer (normalizer.py:46-51: x / s[c]), E3D mu, permute back
    to (B,S,latent,H/4,W/4)."""
    s = torch.tensor(norm_factors, dtype=torch.float32).view(1, -1, 1, 1, 1)
    x = velocity_3d.permute(0, 2, 1, 3, 4) / s
    mu, _ = ovae.encoder_forward(vae_sd, x, "encoder_3d.")
    return mu.permute(0, 2, 1, 3, 4)


def decode(vae_sd, x, B, img, norm_factors):
    """predictor.py:993-1021: reshape, D3D decode, denormalise, mask."""
    N, lc, lh, lw = x.shape
    ld = N // B
    z = x.reshape(B, ld, lc, lh, lw).permute(0, 2, 1, 3, 4)
    v = ovae.decode_3d(vae_sd, z).permute(0, 2, 1, 3, 4)  # (B,S,3,H,W)
    s = torch.tensor(norm_factors, dtype=torch.float32).view(1, 1, -1, 1, 1)
    return v * s * img


def predict_ddim(unet_sd, vae_sd, img, velocity_2d, noise, num_steps=50, eta=0.0, norm_factors=(1, 1, 1),
                 num_timesteps=1000, use_edt=True, record=None, step_noise=None):
    """predictor.py:898-1023.  `record` (list) receives (t, x_t, eps, x_prev) per step."""
    B = img.shape[0]
    sch = OracleScheduler(num_timesteps)
    v_lat, feats = conditioning(vae_sd, img, velocity_2d, norm_factors, use_edt)
    x = noise.reshape(v_lat.shape[0], v_lat.shape[1], v_lat.shape[2], v_lat.shape[3])
    ts = ddim_timesteps(num_timesteps, num_steps)
    for i, t in enumerate(ts):
        t_prev = ts[i + 1] if i + 1 < len(ts) else -1
        tb = torch.full((x.shape[0],), t, dtype=torch.long)
        eps = ounet.unet_forward(unet_sd, torch.cat([x, v_lat, feats], 1), tb)
        z = None if step_noise is None else step_noise[i]
        x_new = sch.ddim_sample(eps, x, t, t_prev, eta=eta, clip_range=(-30.0, 30.0), noise=z)
        if record is not None:
            record.append((t, x, eps, x_new))
        x = x_new
    return decode(vae_sd, x, B, img, norm_factors)


def predict(unet_sd, vae_sd, img, velocity_2d, noise, step_noise, norm_factors=(1, 1, 1), num_timesteps=1000,
            use_edt=True, record=None):
    """predictor.py:754-896 (multi-step branch :841-851).  step_noise[i] is the i-th
    torch.randn_like draw of p_sample (diffusion.py:175), i.e. for t = T-1-i."""
    B = img.shape[0]
    sch = OracleScheduler(num_timesteps)
    v_lat, feats = conditioning(vae_sd, img, velocity_2d, norm_factors, use_edt)
    x = noise.reshape(v_lat.shape)
    if num_timesteps == 1:  # predictor.py:823-838
        tb = torch.zeros((x.shape[0],), dtype=torch.long)
        eps = ounet.unet_forward(unet_sd, torch.cat([x, v_lat, feats], 1), tb)
        ab = sch.alphas_cumprod[0]
        x = torch.clamp((x - torch.sqrt(1 - ab) * eps) / torch.sqrt(ab), -30.0, 30.0)
    else:
        for i, t in enumerate(reversed(range(num_timesteps))):
            tb = torch.full((x.shape[0],), t, dtype=torch.long)
            eps = ounet.unet_forward(unet_sd, torch.cat([x, v_lat, feats], 1), tb)
            x_new = sch.p_sample(eps, x, t, step_noise[i], clip_denoised=True, clip_range=(-30.0, 30.0))
            if record is not None:
                record.append((t, x, eps, x_new))
            x = x_new
    return decode(vae_sd, x, B, img, norm_factors)
