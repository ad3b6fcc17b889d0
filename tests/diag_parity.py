"""Where does the bf16 path's end-to-end error come from?  One sample of 11x256x256, DDIM-N: the GPU path against the CPU
oracle, stage by stage -- conditioning (E2D mu, distance features), the latent after the loop, the decode alone (D3D of the
ORACLE's final latent), and the loop alone (GPU loop started from the oracle's conditioning).
usage: python tests/diag_parity.py [steps=50] [precision=bf16] [B=1]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from diffusion_model_project_b200 import _lib, engine, synth  # noqa: E402
from diffusion_model_project_b200.predictor import B200LatentDiffusionPredictor  # noqa: E402
from oracle import predictor as opred  # noqa: E402
from util import rel_err, rel_l2  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
precision = sys.argv[2] if len(sys.argv) > 2 else "f16"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
S, SIZE = 11, 256
torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count() or 1)
usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
img, v2d = synth.synth_inputs(B, num_slices=S, size=SIZE, seed=2024)
noise = synth.synth_noise(B, num_slices=S, latent_size=SIZE // 4, seed=42)

rec_ref = []
ref = opred.predict_ddim(usd, vsd, img, v2d, noise, num_steps=steps, eta=0.0, norm_factors=synth.NORM_FACTORS, record=rec_ref)
v_ref, f_ref = opred.conditioning(vsd, img, v2d, synth.NORM_FACTORS)
x_ref = rec_ref[-1][3]

p = B200LatentDiffusionPredictor("UNet", dict(synth.UNET_KWARGS), True, unet_state=usd, vae_state=vsd, norm_factors=synth.NORM_FACTORS,
                                 num_slices=S, num_timesteps=1000, precision=precision, use_graph=False, device="cuda")
rec = []
out = p.predict_ddim(img.cuda(), v2d.cuda(), num_steps=steps, eta=0.0, noise=noise.cuda(), record=rec).cpu()
ses = p._session
ui = ses["unet_in"]
c = (ui.hi.view(torch.float16).float() if ui.f16 else ui.hi.float()) if ui.lo is None else ui.hi.float() + ui.lo.float()
c = c[:, 0, :, :, 8:17].permute(0, 3, 1, 2).contiguous().cpu()
print(f"[{precision}, DDIM-{steps}, B={B}] end-to-end field rel-L2 {rel_l2(out, ref):.3e}  max-rel {rel_err(out, ref):.3e}")
print(f"  conditioning: E2D mu rel-L2 {rel_l2(c[:, :8], v_ref):.3e} max-rel {rel_err(c[:, :8], v_ref):.3e}; features rel-L2 {rel_l2(c[:, 8:9], f_ref):.3e}")
x_gpu = ses["x"].permute(0, 3, 1, 2).contiguous().cpu()
print(f"  latent after the loop: rel-L2 {rel_l2(x_gpu, x_ref):.3e} max-rel {rel_err(x_gpu, x_ref):.3e}   (|x| rms {x_ref.pow(2).mean().sqrt():.3f})")
for i in sorted(set(list(range(0, steps, max(1, steps // 10))) + [steps - 1])):
    print(f"    step {i:3d}: x_t rel-L2 {rel_l2(rec[i][0].cpu(), rec_ref[i][1]):.3e}  eps (own trajectories) rel-L2 {rel_l2(rec[i][1].cpu(), rec_ref[i][2]):.3e}"
          f"  max-rel {rel_err(rec[i][1].cpu(), rec_ref[i][2]):.3e}")
# decode alone: D3D of the oracle's final latent (through the module API), denormalised and masked like the predictor does
z = x_ref.reshape(B, S, 8, SIZE // 4, SIZE // 4).permute(0, 2, 1, 3, 4).contiguous()
dec = p.vae.decode_3d(z.cuda()).permute(0, 2, 1, 3, 4).cpu()
dec = dec * torch.tensor(synth.NORM_FACTORS).view(1, 1, 3, 1, 1) * img
print(f"  decode alone (D3D of the oracle's latent): rel-L2 {rel_l2(dec, ref):.3e} max-rel {rel_err(dec, ref):.3e}")
# loop alone: the GPU loop started from the ORACLE's conditioning
s = _lib.stream_ptr()
cond = torch.cat([v_ref, f_ref], 1).permute(0, 2, 3, 1).contiguous().cuda()   # (N, h, w, 9)
N = cond.shape[0]
engine.planar_to_cl(cond, ui, N * cond.shape[1] * cond.shape[2], 9, 1, 8, None, s)
p._set_latent(ses, noise.cuda(), s)
p._run_loop(ses, 1, ses["coef"], steps, None, (-30.0, 30.0), [], False)
x_gpu2 = ses["x"].permute(0, 3, 1, 2).contiguous().cpu()
print(f"  loop alone (oracle conditioning, rounded to the path's storage): latent rel-L2 {rel_l2(x_gpu2, x_ref):.3e} max-rel {rel_err(x_gpu2, x_ref):.3e}")
out2 = p._decode(ses, s).cpu()
print(f"  loop + decode from the oracle's conditioning: field rel-L2 {rel_l2(out2, ref):.3e}")
