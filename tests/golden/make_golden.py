"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference).

Run once in the build container (the reference tree is not present on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (committed): tests/golden/reference_keys.json, scheduler.npz, unet.npz, vae.npz,
predict_ddim.npz, predict_ddpm.npz.  Weights and inputs are NOT stored: both sides regenerate
them from diffusion_model_project_b200.synth (seeded, bit-identical on CPU).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.environ.get("B2D_GOLDEN_OUT", HERE)  # where the vectors are written (tests/test_golden_regen.py redirects it)
sys.path.insert(0, ROOT)
sys.path[:0] = ["/root/reference", "/root/reference/Diffusion_model"]
os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True

from diffusion_model_project_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)
torch.manual_seed(0)


def build_reference_predictor(tmp, num_timesteps=1000, num_slices=11):
    from src.predictor import LatentDiffusionPredictor
    from VAE_model.src.dual_vae.model import DualBranchVAE

    vae = DualBranchVAE(3, 8)
    vsd = synth.synth_vae_state(seed=1)
    vae.encoder_2d.load_state_dict({k[len("encoder_2d."):]: v for k, v in vsd.items() if k.startswith("encoder_2d.")})
    vae.decoder_3d.load_state_dict({k[len("decoder_3d."):]: v for k, v in vsd.items() if k.startswith("decoder_3d.")})
    for d in ("enc", "dec"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
        torch.save(vae.state_dict(), os.path.join(tmp, d, "model.pt"))
        with open(os.path.join(tmp, d, "vae_log.json"), "w") as f:
            json.dump({"norm_factors": synth.NORM_FACTORS, "latent_channels": 8, "in_channels": 3}, f)
    kw = dict(synth.UNET_KWARGS)
    kw.pop("time_embedding_dim")  # injected by the predictor (predictor.py:317-318)
    pred = LatentDiffusionPredictor(
        "UNet", model_kwargs=kw, distance_transform=True,
        vae_encoder_path=os.path.join(tmp, "enc"), vae_decoder_path=os.path.join(tmp, "dec"),
        num_slices=num_slices, num_timesteps=num_timesteps,
    )
    pred.model.load_state_dict(synth.synth_unet_state(seed=0))
    pred.eval()
    pred.to("cpu")
    return pred


class NoiseFeeder:
    """Serve host-injected noise in place of torch.randn_like (diffusion.py:175,231)."""

    def __init__(self, draws):
        self.draws = list(draws)
        self.i = 0
        self._orig = torch.randn_like

    def __enter__(self):
        def fake(x, *a, **k):
            z = self.draws[self.i].reshape(x.shape).to(x.dtype)
            self.i += 1
            return z
        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def main():
    from src.unet.models import UNet
    from src.diffusion import DiffusionScheduler
    from VAE_model.src.dual_vae.model import DualBranchVAE

    # ---- 1. key / shape lists --------------------------------------------------------------
    unet = UNet(**synth.UNET_KWARGS)
    vae = DualBranchVAE(3, 8)
    keys = {
        "unet": [[k, list(v.shape)] for k, v in unet.state_dict().items()],
        "vae": [[k, list(v.shape)] for k, v in vae.state_dict().items()],
    }
    with open(os.path.join(OUT, "reference_keys.json"), "w") as f:
        json.dump(keys, f)

    # ---- 2. scheduler -----------------------------------------------------------------------
    sch = DiffusionScheduler(1000, device="cpu")
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 8, 16, 16, generator=g) * 3
    eps = torch.randn(2, 8, 16, 16, generator=g)
    z = torch.randn(2, 8, 16, 16, generator=g)
    out = {k: getattr(sch, k).numpy() for k in (
        "betas", "alphas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
        "sqrt_one_minus_alphas_cumprod", "posterior_variance", "posterior_log_variance",
        "posterior_mean_coef1", "posterior_mean_coef2")}
    out.update(x=x.numpy(), eps=eps.numpy(), z=z.numpy())
    ts = [999, 998, 500, 37, 1, 0]
    out["p_ts"] = np.array(ts)
    for t in ts:
        with NoiseFeeder([z]):
            out[f"p_sample_{t}"] = sch.p_sample(eps, x, t, clip_denoised=True, clip_range=(-30.0, 30.0)).numpy()
        out[f"q_sample_{t}"] = sch.q_sample(x, t, eps).numpy()
        out[f"x0_{t}"] = sch.predict_x0_from_noise(x, t, eps).numpy()
    with NoiseFeeder([z]):
        out["p_sample_default_clip_999"] = sch.p_sample(eps, x, 999).numpy()
    pairs = [(999, 978), (978, 958), (509, 489), (20, 0), (0, -1)]
    out["ddim_pairs"] = np.array(pairs)
    for t, tp in pairs:
        out[f"ddim_{t}_{tp}"] = sch.ddim_sample(eps, x, t, tp, eta=0.0).numpy()
        with NoiseFeeder([z]):
            out[f"ddim_eta05_{t}_{tp}"] = sch.ddim_sample(eps, x, t, tp, eta=0.5).numpy()
    np.savez_compressed(os.path.join(OUT, "scheduler.npz"), **out)

    # ---- 3. UNet forward ---------------------------------------------------------------------
    unet.load_state_dict(synth.synth_unet_state(seed=0))
    unet.eval()
    g = torch.Generator().manual_seed(11)
    xin = torch.randn(2, 17, 32, 32, generator=g)
    tt = torch.tensor([999, 500], dtype=torch.long)
    eps_out = unet(xin, tt)
    np.savez_compressed(os.path.join(OUT, "unet.npz"), eps=eps_out.numpy(), t=tt.numpy())
    print("unet eps", eps_out.abs().max().item(), eps_out.std().item())

    # ---- 4. VAE branches -----------------------------------------------------------------------
    vsd = synth.synth_vae_state(seed=1)
    vae.encoder_2d.load_state_dict({k[len("encoder_2d."):]: v for k, v in vsd.items() if k.startswith("encoder_2d.")})
    vae.decoder_3d.load_state_dict({k[len("decoder_3d."):]: v for k, v in vsd.items() if k.startswith("decoder_3d.")})
    vae.eval()
    g = torch.Generator().manual_seed(13)
    xv = torch.randn(1, 3, 3, 32, 32, generator=g)
    zz, (mu, logvar) = vae.encode_2d_deterministic(xv)
    zl = torch.randn(1, 8, 3, 8, 8, generator=g)
    dec = vae.decode_3d(zl)
    np.savez_compressed(os.path.join(OUT, "vae.npz"), mu=mu.numpy(), logvar=logvar.numpy(), dec=dec.numpy())
    print("vae mu", mu.abs().max().item(), "dec", dec.abs().max().item())

    # ---- 5. predict_ddim / predict end to end (small) --------------------------------------------
    with tempfile.TemporaryDirectory() as tmp:
        pred = build_reference_predictor(tmp, num_timesteps=1000, num_slices=2)
        img, v2d = synth.synth_inputs(1, num_slices=2, size=128, seed=2024)
        noise = synth.synth_noise(1, num_slices=2, latent_size=32, seed=42)
        rec = []
        orig = pred.model.forward

        def spy(xx, t):
            e = orig(xx, t)
            rec.append(e.clone())
            return e
        pred.model.forward = spy
        out_ddim = pred.predict_ddim(img, v2d, num_steps=3, eta=0.0, noise=noise.clone())
        eps_steps = torch.stack(rec)
        rec.clear()
        g = torch.Generator().manual_seed(99)
        zs = [torch.randn(2, 8, 32, 32, generator=g) for _ in range(3)]
        with NoiseFeeder(zs):
            out_ddim_eta = pred.predict_ddim(img, v2d, num_steps=3, eta=0.7, noise=noise.clone())
        rec.clear()
        np.savez_compressed(os.path.join(OUT, "predict_ddim.npz"), out=out_ddim.numpy(), eps_steps=eps_steps.numpy(),
                            out_eta07=out_ddim_eta.numpy())
        print("ddim out", out_ddim.abs().max().item())

    with tempfile.TemporaryDirectory() as tmp:
        pred = build_reference_predictor(tmp, num_timesteps=12, num_slices=2)
        g = torch.Generator().manual_seed(100)
        zs = [torch.randn(2, 8, 32, 32, generator=g) for _ in range(12)]
        with NoiseFeeder(zs):
            out_ddpm = pred.predict(img, v2d, noise=noise.clone())
        np.savez_compressed(os.path.join(OUT, "predict_ddpm.npz"), out=out_ddpm.numpy())
        print("ddpm out", out_ddpm.abs().max().item())


if __name__ == "__main__":
    main()
