"""B200Scheduler -- drop-in for DiffusionScheduler (Diffusion_model/src/diffusion.py:33-234).

Same constructor, same ten buffers as attributes, same method names and argument meaning.
`p_sample` / `ddim_sample` / `predict_x0_from_noise` / `q_sample` each launch ONE fused libb2d
kernel (csrc/scheduler.cu) instead of ~10 ATen elementwise kernels.  Coefficients are looked up
from a device table row, so the same kernel serves the graph-captured loop of the predictor.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib


class B200Scheduler:
    def __init__(self, num_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02, device="cuda"):
        self.num_timesteps = num_timesteps
        self.device = torch.device(device)
        # diffusion.py:45-76 -- float64 tables cast to float32 (host, exact same ops as the reference)
        betas = torch.linspace(beta_start, beta_end, num_timesteps, dtype=torch.float64)
        alphas = 1.0 - betas
        ac = torch.cumprod(alphas, dim=0)
        ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
        pv = torch.clamp(betas * (1.0 - ac_prev) / (1.0 - ac), min=1e-20)
        host = {
            "betas": betas.float(), "alphas": alphas.float(), "alphas_cumprod": ac.float(),
            "alphas_cumprod_prev": ac_prev.float(), "sqrt_alphas_cumprod": torch.sqrt(ac).float(),
            "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac).float(), "posterior_variance": pv.float(),
            "posterior_log_variance": torch.log(pv).float(),
            "posterior_mean_coef1": (betas * torch.sqrt(ac_prev) / (1.0 - ac)).float(),
            "posterior_mean_coef2": ((1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac)).float(),
        }
        self._host = host
        self._names = tuple(host.keys())
        self._ddim_rows: Dict[Tuple[int, int, float], torch.Tensor] = {}
        self._place()

    def _place(self):
        for k, v in self._host.items():
            setattr(self, k, v.to(self.device))
        self._ddpm_table = self.ddpm_coef_rows(range(self.num_timesteps)).to(self.device)
        self._x0_table = self.x0_coef_rows().to(self.device)
        self._ddim_rows.clear()

    def to(self, device):
        self.device = torch.device(device)
        self._place()
        return self

    # ------------------------------------------------------------------ coefficient rows {a,b,c1,c2,s,0,0,0}
    def ddpm_coef_rows(self, timesteps) -> torch.Tensor:
        """p_sample coefficients for each t (diffusion.py:119-123,146-148,178-181)."""
        h = self._host
        rows = []
        for t in timesteps:
            a = torch.clamp(h["sqrt_alphas_cumprod"][t], min=1e-8)
            s = torch.sqrt(h["posterior_variance"][t]) if t != 0 else torch.tensor(0.0)
            rows.append(torch.stack([a, h["sqrt_one_minus_alphas_cumprod"][t], h["posterior_mean_coef1"][t],
                                     h["posterior_mean_coef2"][t], s, *([torch.tensor(0.0)] * 3)]))
        return torch.stack(rows).float().contiguous()

    def x0_coef_rows(self) -> torch.Tensor:
        h = self._host
        T = self.num_timesteps
        rows = torch.zeros(T, 8)
        rows[:, 0] = torch.clamp(h["sqrt_alphas_cumprod"], min=1e-8)
        rows[:, 1] = h["sqrt_one_minus_alphas_cumprod"]
        rows[:, 2] = 1.0
        return rows.contiguous()

    def ddim_coef_row(self, t: int, t_prev: int, eta: float) -> torch.Tensor:
        """ddim_sample coefficients, computed with the reference's fp32 expression order (diffusion.py:208-232)."""
        h = self._host
        ab_t = h["alphas_cumprod"][t]
        ab_p = h["alphas_cumprod"][t_prev] if t_prev >= 0 else torch.tensor(1.0)
        a = torch.clamp(h["sqrt_alphas_cumprod"][t], min=1e-8)
        b = h["sqrt_one_minus_alphas_cumprod"][t]
        sigma = eta * torch.sqrt((1 - ab_p) / (1 - ab_t) * (1 - ab_t / ab_p))
        c2 = torch.sqrt(1 - ab_p - sigma ** 2)
        c1 = torch.sqrt(ab_p)
        s = sigma if (eta > 0 and t > 0) else torch.tensor(0.0)
        z = torch.tensor(0.0)
        return torch.stack([a, b, c1, c2, s.float() if torch.is_tensor(s) else torch.tensor(float(s)), z, z, z]).float()

    def ddim_coef_rows(self, timesteps, eta: float) -> torch.Tensor:
        ts = list(timesteps)
        return torch.stack([self.ddim_coef_row(t, ts[i + 1] if i + 1 < len(ts) else -1, eta) for i, t in enumerate(ts)]).contiguous()

    # ------------------------------------------------------------------ kernel launch
    def _step(self, kind, x_t, eps, noise, row_table, row, clip, clip_range, out=None):
        if not x_t.is_cuda:
            raise RuntimeError("B200Scheduler runs on a CUDA device only (no CPU fallback)")
        x_t = x_t.contiguous().float()
        eps = eps.contiguous().float()
        if noise is not None:
            noise = noise.contiguous().float()
        if out is None:
            out = torch.empty_like(x_t)
        # every caller below passes a noise tensor wherever the row's s != 0 (drawn from torch's generator like the
        # reference's randn_like), so the lean kernel without the Philox generator is selected (seed 0)
        _lib.call("b2d_scheduler_step", kind, x_t.data_ptr(), eps.data_ptr(), _lib.ptr(noise), out.data_ptr(), x_t.numel(),
                  row_table.data_ptr(), None, int(row), 0, 1 if clip else 0, float(clip_range[0]), float(clip_range[1]),
                  None, 0, 0, 0, None, None, 0, _lib.stream_ptr())
        return out

    @staticmethod
    def _uniform_t(t):
        if isinstance(t, int):
            return t
        if torch.is_tensor(t):
            if t.dim() == 0:
                return int(t.item())
            vals = t.detach().to("cpu").reshape(-1)
            if bool((vals == vals[0]).all()):
                return int(vals[0].item())
            return None
        return int(t)

    def _per_image(self, t, x, fn):
        """Tensor t with different timesteps per sample: one launch per image (not on the sampling path)."""
        tl = t.detach().to("cpu").reshape(-1).tolist()
        return torch.stack([fn(int(tv), x[i:i + 1], i)[0] for i, tv in enumerate(tl)])

    # ------------------------------------------------------------------ reference API
    def q_sample(self, x_start, t, noise=None):
        """diffusion.py:78-101."""
        if noise is None:
            noise = torch.randn_like(x_start)
        n = x_start.shape[0]
        if isinstance(t, int) or (torch.is_tensor(t) and t.dim() == 0):
            t = torch.full((n,), int(t), dtype=torch.long, device=x_start.device)
        t = t.to(self.device)
        a = self.sqrt_alphas_cumprod[t].contiguous()
        b = self.sqrt_one_minus_alphas_cumprod[t].contiguous()
        x_start = x_start.contiguous().float()
        noise = noise.contiguous().float()
        out = torch.empty_like(x_start)
        _lib.call("b2d_q_sample", x_start.data_ptr(), noise.data_ptr(), out.data_ptr(), a.data_ptr(), b.data_ptr(), n,
                  x_start.numel() // n, _lib.stream_ptr())
        return out

    def predict_x0_from_noise(self, x_t, t, noise):
        """diffusion.py:103-125."""
        tu = self._uniform_t(t)
        if tu is None:
            return self._per_image(t, x_t, lambda tv, xs, i: self._step(0, xs, noise[i:i + 1], None, self._x0_table, tv, False, (0, 0)))
        return self._step(0, x_t, noise, None, self._x0_table, tu, False, (0.0, 0.0))

    def q_posterior_mean_variance(self, x_0, x_t, t):
        """diffusion.py:127-150 (helper, not on the sampling path: p_sample fuses it)."""
        c1, c2, v = self.posterior_mean_coef1[t], self.posterior_mean_coef2[t], self.posterior_variance[t]
        while c1.dim() < x_0.dim():
            c1, c2, v = c1.unsqueeze(-1), c2.unsqueeze(-1), v.unsqueeze(-1)
        return c1 * x_0 + c2 * x_t, v

    def p_sample(self, model_output, x_t, t, clip_denoised=True, clip_range=(-20.0, 20.0), noise=None):
        """diffusion.py:152-188.  `noise` replaces the reference's torch.randn_like(x_t) draw (taken
        from torch's generator when None, one draw per call like the reference, even at t == 0)."""
        if noise is None:
            noise = torch.randn_like(x_t)
        tu = self._uniform_t(t)
        if tu is None:
            return self._per_image(t, x_t, lambda tv, xs, i: self._step(0, xs, model_output[i:i + 1], noise[i:i + 1],
                                                                       self._ddpm_table, tv, clip_denoised, clip_range))
        return self._step(0, x_t, model_output, noise, self._ddpm_table, tu, clip_denoised, clip_range)

    def ddim_sample(self, model_output, x_t, t, t_prev, eta=0.0, clip_range=(-30.0, 30.0), noise=None):
        """diffusion.py:195-234."""
        t, t_prev = int(t), int(t_prev)
        key = (t, t_prev, float(eta))
        row = self._ddim_rows.get(key)
        if row is None:
            row = self.ddim_coef_row(t, t_prev, eta).to(self.device)
            if len(self._ddim_rows) > 4096:
                self._ddim_rows.clear()
            self._ddim_rows[key] = row
        if eta > 0 and t > 0 and noise is None:
            noise = torch.randn_like(x_t)
        return self._step(1, x_t, model_output, noise, row, 0, True, clip_range)
