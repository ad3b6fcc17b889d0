"""Error attribution of the bf16 path on the golden predict_ddim case (1 x 2 x 128 x 128, 3 DDIM steps):
per-stage errors against the CPU oracle.  usage: python tests/diag_precision.py [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from diffusion_model_project_b200 import synth  # noqa: E402
from diffusion_model_project_b200.predictor import B200LatentDiffusionPredictor  # noqa: E402
from oracle import predictor as opred, vae as ovae  # noqa: E402
from util import rel_err, rel_l2  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
torch.set_grad_enabled(False)
usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
img, v2d = synth.synth_inputs(1, num_slices=2, size=128, seed=2024)
noise = synth.synth_noise(1, num_slices=2, latent_size=32, seed=42)
rec_o = []
ref = opred.predict_ddim(usd, vsd, img, v2d, noise, num_steps=3, eta=0.0, norm_factors=synth.NORM_FACTORS, record=rec_o)
p = B200LatentDiffusionPredictor("UNet", dict(synth.UNET_KWARGS), True, unet_state=usd, vae_state=vsd, norm_factors=synth.NORM_FACTORS,
                                 num_slices=2, num_timesteps=1000, precision=prec, use_graph=False, device="cuda")
rec = []
out = p.predict_ddim(img.cuda(), v2d.cuda(), num_steps=3, eta=0.0, noise=noise.cuda(), record=rec).cpu()
print(f"[{prec}] final field rel-L2 = {rel_l2(out, ref):.4e}  max-rel = {rel_err(out, ref):.4e}")
for i in range(3):
    t, x_o, eps_o, xn_o = rec_o[i]
    x_b, eps_b, xn_b = (r.cpu() for r in rec[i])
    print(f"  step {i} t={t}: x_t rel-L2 {rel_l2(x_b, x_o):.3e} | eps max-rel {rel_err(eps_b, eps_o):.3e} rel-L2 {rel_l2(eps_b, eps_o):.3e} | x_next rel-L2 {rel_l2(xn_b, xn_o):.3e}")
# conditioning (E2D mu) error: channels 8..15 of the UNet input buffer
ses = p._session
v_lat_o, feats_o = opred.conditioning(vsd, img, v2d, synth.NORM_FACTORS)
ui = ses["unet_in"].hi.float()
if ses["unet_in"].lo is not None:
    ui = ui + ses["unet_in"].lo.float()
mu_b = ui[:, 0, :, :, 8:16].permute(0, 3, 1, 2).cpu()
ft_b = ui[:, 0, :, :, 16].cpu()
print(f"  E2D mu rel-L2 {rel_l2(mu_b, v_lat_o):.3e} max-rel {rel_err(mu_b, v_lat_o):.3e}; feats rel-L2 {rel_l2(ft_b, feats_o[:, 0]):.3e}")
# decoder alone on the oracle's final latent
x_fin = rec_o[-1][3]
z = x_fin.reshape(1, 2, 8, 32, 32).permute(0, 2, 1, 3, 4).contiguous()
dec_b = p.vae.decode_3d(z.cuda()).cpu()
dec_o = ovae.decode_3d(vsd, z)
print(f"  D3D alone on the oracle latent: rel-L2 {rel_l2(dec_b, dec_o):.3e} max-rel {rel_err(dec_b, dec_o):.3e}")
# decoder on the GPU's final latent, decoded by the oracle -> error due to the latent only
x_gpu = rec[-1][2].cpu()
ref_from_gpu_latent = opred.decode(vsd, x_gpu, 1, img, synth.NORM_FACTORS)
print(f"  oracle decode of the GPU latent vs reference: rel-L2 {rel_l2(ref_from_gpu_latent, ref):.3e}  (latent rel-L2 {rel_l2(x_gpu, x_fin):.3e})")
