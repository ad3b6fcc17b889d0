"""Oracle restatement of one UNet training step (SURVEY.md section 8, row f4 -- the next row after the sampling path).

Test infrastructure (see oracle/__init__.py); nothing in the product package imports it.  Follows
  * Diffusion_model/src/predictor.py:722-748  -- q_sample of the target latents, concat with the conditioning, eps-prediction
  * Diffusion_model/src/unet/metrics.py:337-402 -- normalized_mse_loss_per_component (train.py:150,155 default criterion)
  * Diffusion_model/src/helper.py:428-430      -- zero_grad / backward / optimizer.step
  * Diffusion_model/train.py:144-148           -- torch.optim.Adam(lr, weight_decay), defaults betas (0.9, 0.999), eps 1e-8
The backward pass is plain autograd through the functional oracle UNet (oracle/unet.py); the Adam update is restated.
Pinned by tests/golden/train_step.npz (tests/golden/make_train_golden.py runs the unmodified reference modules).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import unet as ounet
from .scheduler import OracleScheduler


def normalized_mse_loss_per_component(output: torch.Tensor, target: torch.Tensor, reduce: bool = True,
                                      weight_per_channel: Optional[torch.Tensor] = None, eps: float = 1e-8) -> torch.Tensor:
    """metrics.py:337-402: per (sample, channel) MSE over the spatial dims divided by the target's mean square (+eps),
    optionally channel-weighted, averaged over channels and (reduce) over the batch."""
    if output.dim() == 4:
        dims = (-2, -1)
    elif output.dim() == 5:
        dims = (-3, -2, -1)
    else:
        raise ValueError(f"Expected 4D or 5D tensor, got {output.dim()}D")
    mse = torch.mean((output - target) ** 2, dim=dims)
    norm = torch.mean(target ** 2, dim=dims)
    err = mse / (norm + eps)
    if weight_per_channel is not None:
        w = weight_per_channel.unsqueeze(0) if weight_per_channel.dim() == 1 else weight_per_channel
        err = err * w / w.sum()
    err = torch.mean(err, dim=-1)
    return err.mean() if reduce else err


def training_loss_and_grads(sd: Dict[str, torch.Tensor], x_start: torch.Tensor, cond: torch.Tensor, feats: torch.Tensor,
                            t: torch.Tensor, noise: torch.Tensor, num_timesteps: int = 1000
                            ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]:
    """predictor.py:722-748 + helper.py:428-429.  x_start, noise: (N, 8, h, w) target latents / injected noise; cond: (N, 8, h, w)
    E2D latents; feats: (N, 1, h, w) mask features; t: (N,) long.  Returns (loss, {param name: grad}, noise_pred)."""
    sch = OracleScheduler(num_timesteps)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        x_t = sch.q_sample(x_start, t, noise)
        unet_in = torch.cat([x_t, cond, feats], dim=1)
        pred = ounet.unet_forward(params, unet_in, t)
        loss = normalized_mse_loss_per_component(pred, noise)
        names = list(params.keys())
        grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    out = {k: (g.detach() if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}
    return loss.detach(), out, pred.detach()


def training_step_from_fields(unet_sd: Dict[str, torch.Tensor], vae_sd: Dict[str, torch.Tensor], img: torch.Tensor,
                              velocity_2d: torch.Tensor, velocity_3d: torch.Tensor, t: torch.Tensor, noise: torch.Tensor,
                              norm_factors, num_timesteps: int = 1000):
    """The 'latent-diffusion' branch of the training loop body, helper.py:277-430 with its default losses, from the fields:
    target latents = encode_target(velocity_3d) (predictor.py:1042-1085, helper.py:288), then predictor.forward
    (predictor.py:636-751): frozen E2D mu of the normalised 2D velocity + EDT features as conditioning, q_sample at the
    given timesteps (the reference draws them with randint, :736 -- injected here), UNet, criterion, backward.
    noise: (B, S, 8, h, w) like helper.py:299.  Returns (loss, {param: grad}, noise_pred, (x_start, cond, feats))."""
    from . import predictor as opred
    with torch.no_grad():
        latents = opred.encode_target(vae_sd, velocity_3d, norm_factors)           # (B, S, 8, h, w)
        cond, feats = opred.conditioning(vae_sd, img, velocity_2d, norm_factors)    # (N, 8, h, w), (N, 1, h, w)
    N = cond.shape[0]
    x_start = latents.reshape(N, *latents.shape[2:])
    noise_flat = noise.reshape(x_start.shape)
    loss, grads, pred = training_loss_and_grads(unet_sd, x_start, cond, feats, t, noise_flat, num_timesteps)
    return loss, grads, pred, (x_start, cond, feats)


def adam_step(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float = 1e-4,
              betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
    """torch.optim.Adam (train.py:144-148), single-tensor form: L2 weight decay folded into the gradient, bias-corrected
    moments, update = lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  Returns (param, m, v)."""
    b1, b2 = betas
    if weight_decay != 0.0:
        grad = grad + weight_decay * param
    m = b1 * m + (1.0 - b1) * grad
    v = b2 * v + (1.0 - b2) * grad * grad
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + eps
    return param - (lr / bc1) * m / denom, m, v
