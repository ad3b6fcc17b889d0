"""Oracle restatement of DiffusionScheduler (Diffusion_model/src/diffusion.py:33-234).

Test infrastructure (see oracle/__init__.py).  Tables are built in float64 and cast
to float32 exactly as diffusion.py:45-76; all step arithmetic is fp32 on CPU.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class OracleScheduler:
    def __init__(self, num_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02):
        # diffusion.py:45-50
        self.num_timesteps = num_timesteps
        betas = torch.linspace(beta_start, beta_end, num_timesteps, dtype=torch.float64)
        alphas = 1.0 - betas
        ac = torch.cumprod(alphas, dim=0)
        ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
        # diffusion.py:53-76
        self.betas = betas.float()
        self.alphas = alphas.float()
        self.alphas_cumprod = ac.float()
        self.alphas_cumprod_prev = ac_prev.float()
        self.sqrt_alphas_cumprod = torch.sqrt(ac).float()
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - ac).float()
        pv = torch.clamp(betas * (1.0 - ac_prev) / (1.0 - ac), min=1e-20)
        self.posterior_variance = pv.float()
        self.posterior_log_variance = torch.log(pv).float()
        self.posterior_mean_coef1 = (betas * torch.sqrt(ac_prev) / (1.0 - ac)).float()
        self.posterior_mean_coef2 = ((1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac)).float()

    # diffusion.py:78-101
    def q_sample(self, x_start, t, noise):
        a = self.sqrt_alphas_cumprod[t]
        b = self.sqrt_one_minus_alphas_cumprod[t]
        while a.dim() < x_start.dim():
            a = a.unsqueeze(-1)
            b = b.unsqueeze(-1)
        return a * x_start + b * noise

    # diffusion.py:103-125
    def predict_x0_from_noise(self, x_t, t, noise):
        a = self.sqrt_alphas_cumprod[t]
        b = self.sqrt_one_minus_alphas_cumprod[t]
        while a.dim() < x_t.dim():
            a = a.unsqueeze(-1)
            b = b.unsqueeze(-1)
        a = torch.clamp(a, min=1e-8)
        return (x_t - b * noise) / a

    # diffusion.py:152-188 ; `noise` replaces the reference's torch.randn_like(x_t) draw
    def p_sample(self, model_output, x_t, t: int, noise, clip_denoised=True, clip_range=(-20.0, 20.0)):
        x0 = self.predict_x0_from_noise(x_t, t, model_output)
        if clip_denoised:
            x0 = torch.clamp(x0, clip_range[0], clip_range[1])
        mean = self.posterior_mean_coef1[t] * x0 + self.posterior_mean_coef2[t] * x_t
        if t == 0:
            return mean
        return mean + torch.sqrt(self.posterior_variance[t]) * noise

    # diffusion.py:195-234
    def ddim_sample(self, model_output, x_t, t: int, t_prev: int, eta: float = 0.0, clip_range=(-30.0, 30.0), noise=None):
        ab_t = self.alphas_cumprod[t]
        ab_p = self.alphas_cumprod[t_prev] if t_prev >= 0 else torch.tensor(1.0)
        x0 = self.predict_x0_from_noise(x_t, t, model_output)
        x0 = torch.clamp(x0, clip_range[0], clip_range[1])
        sigma = eta * torch.sqrt((1 - ab_p) / (1 - ab_t) * (1 - ab_t / ab_p))
        pred_dir = torch.sqrt(1 - ab_p - sigma ** 2) * model_output
        x_prev = torch.sqrt(ab_p) * x0 + pred_dir
        if eta > 0 and t > 0:
            x_prev = x_prev + sigma * noise
        return x_prev


def ddim_timesteps(num_timesteps: int, num_steps: int):
    """predictor.py:965 -- linspace(T-1, 0, num_steps, dtype=long)."""
    return torch.linspace(num_timesteps - 1, 0, num_steps, dtype=torch.long).tolist()
