"""BASELINE.json configs[1]: full-step DDPM sampling (1000 steps) in the fp32-class mode against the CPU oracle
(= the reference's op sequence), 256x256 in-plane, host-injected noise.  Reports
  * per-step noise-prediction error  max|eps - eps_ref| / max|eps_ref|  with the ORACLE's UNet evaluated on the GPU
    path's own x_t (every --every-th step), north-star bound 1e-3;
  * final velocity-field relative L2 against the oracle's own 1000-step trajectory, bound 1e-2.
usage: python tests/parity_ddpm1000.py [--batch 1] [--steps 1000] [--every 25] [--precision fp32x]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from diffusion_model_project_b200 import synth  # noqa: E402
from diffusion_model_project_b200.predictor import B200LatentDiffusionPredictor  # noqa: E402
from oracle import predictor as opred, unet as ounet  # noqa: E402
from util import rel_err, rel_l2  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--every", type=int, default=25)
ap.add_argument("--precision", default="fp32x")
ap.add_argument("--size", type=int, default=256)
a = ap.parse_args()
torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count() or 1)
B, S, T = a.batch, 11, a.steps
usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
img, v2d = synth.synth_inputs(B, num_slices=S, size=a.size, seed=2024)
noise = synth.synth_noise(B, num_slices=S, latent_size=a.size // 4, seed=42)
gen = torch.Generator().manual_seed(100)
zs = [torch.randn(B * S, 8, a.size // 4, a.size // 4, generator=gen) for _ in range(T)]

t0 = time.time()
p = B200LatentDiffusionPredictor("UNet", dict(synth.UNET_KWARGS), True, unet_state=usd, vae_state=vsd, norm_factors=synth.NORM_FACTORS,
                                 num_slices=S, num_timesteps=T, precision=a.precision, use_graph=False, device="cuda")
rec = []
out = p.predict(img.cuda(), v2d.cuda(), noise=noise.cuda(), step_noise=zs, record=rec).cpu()
torch.cuda.synchronize()
print(f"GPU path ({a.precision}): {T} DDPM steps + E2D + D3D in {time.time() - t0:.1f} s (eager, recording every step)", flush=True)

# per-step eps parity on the GPU's own trajectory
v_lat, feats = opred.conditioning(vsd, img, v2d, synth.NORM_FACTORS)
worst = 0.0
for i in range(0, T, a.every):
    x_t, eps_gpu, _ = rec[i]
    t = T - 1 - i
    tb = torch.full((x_t.shape[0],), t, dtype=torch.long)
    eps_ref = ounet.unet_forward(usd, torch.cat([x_t.cpu(), v_lat, feats], 1), tb)
    e = rel_err(eps_gpu.cpu(), eps_ref)
    worst = max(worst, e)
    print(f"  step {i:4d} (t={t:3d}): eps max-rel err {e:.3e}", flush=True)
print(f"per-step eps error, worst of {len(range(0, T, a.every))} sampled steps: {worst:.3e} (bound 1e-3 in fp32 mode, 2e-2 in bf16 mode)", flush=True)

t0 = time.time()
ref = opred.predict(usd, vsd, img, v2d, noise, zs, norm_factors=synth.NORM_FACTORS, num_timesteps=T)
print(f"oracle trajectory on {torch.get_num_threads()} CPU threads: {time.time() - t0:.1f} s", flush=True)
print(f"final velocity field rel-L2 vs the oracle's own {T}-step trajectory: {rel_l2(out, ref):.3e} (bound 1e-2); max-rel {rel_err(out, ref):.3e}")
