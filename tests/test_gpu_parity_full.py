"""Parity at BASELINE.json's own sizes (11 slices of 256x256, batch 2), against the CPU oracle (= the reference's op
sequence, pinned to the reference's outputs by tests/test_oracle_golden.py):

  configs[2]/[3]  16-bit mode, CUDA-graphed DDIM-50 loop   per-step eps <= 2e-2, final field rel-L2 <= 1e-2
  configs[1]      fp32-class mode, full-step DDPM, B = 2   per-step eps <= 1e-3, final field rel-L2 <= 1e-2

The 16-bit mode that meets BOTH bounds at this size is "f16" (IEEE fp16 operands, the default).  With bf16 operands
("bf16" mode) a single UNet evaluation is 1.1e-2 off (inside the 2e-2 per-step bound) but the 50-step trajectory ends
1.8e-2 from the oracle's field -- outside the 1e-2 bound; that measurement is kept below as its own test so the reason
for the default is on record (tests/diag_parity.py decomposes it: E2D 8e-3, loop 1.6e-2, D3D 4e-3).

Per-step eps is the north star's "per-step noise-prediction max relative error": the ORACLE's UNet evaluated on the GPU
path's own UNet input of that step (its x_t, and its E2D / distance conditioning as stored), so the number isolates one
UNet evaluation.  The final field is compared with the oracle's own full trajectory from the same inputs, weights and
host-injected noise.  DDPM-1000: the GPU runs all 1000 steps and eps is checked on every 50th; the full-trajectory
comparison uses num_timesteps = 100 (same code path, same kernels; 1000 oracle UNet steps would take ~25 CPU-minutes).
CPU cost of this file on the GPU box's host cores: about 6 minutes.
"""
import os

import pytest
import torch

from diffusion_model_project_b200 import synth
from diffusion_model_project_b200.predictor import B200LatentDiffusionPredictor
from oracle import predictor as opred
from oracle import unet as ounet
from util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

B, S, SIZE = 2, 11, 256
LAT = SIZE // 4


class StepNoise:
    """step_noise[i]: the i-th per-step N(0,1) draw, regenerated on demand (1000 steps x 2.9 MB stay off the heap)."""

    def __init__(self, seed, n):
        self.seed, self.n = seed, n

    def __getitem__(self, i):
        return torch.randn(B * S, 8, LAT, LAT, generator=torch.Generator().manual_seed(self.seed * 100003 + i))

    def __len__(self):
        return self.n


class EveryNth(list):
    """`record` sink that keeps every n-th step's (x_t, eps, x_next) on the host."""

    def __init__(self, n):
        super().__init__()
        self.n, self.count = n, 0

    def append(self, item):
        if self.count % self.n == 0:
            super().append((self.count,) + tuple(t.cpu() for t in item))
        self.count += 1


@pytest.fixture(scope="module")
def case():
    torch.set_num_threads(os.cpu_count() or 1)
    usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
    img, v2d = synth.synth_inputs(B, num_slices=S, size=SIZE, seed=2024)
    noise = synth.synth_noise(B, num_slices=S, latent_size=LAT, seed=42)
    return usd, vsd, img, v2d, noise


@pytest.fixture(scope="module")
def ddim50_oracle(case):
    """The oracle's own DDIM-50 trajectory (about 40 s of CPU), shared by the two 16-bit-mode tests."""
    usd, vsd, img, v2d, noise = case
    return opred.predict_ddim(usd, vsd, img, v2d, noise, num_steps=50, eta=0.0, norm_factors=synth.NORM_FACTORS)


def _predictor(usd, vsd, precision, T, graph):
    return B200LatentDiffusionPredictor("UNet", dict(synth.UNET_KWARGS), True, unet_state=usd, vae_state=vsd,
                                        norm_factors=synth.NORM_FACTORS, num_slices=S, num_timesteps=T, precision=precision,
                                        use_graph=graph, device="cuda")


def _gpu_conditioning(p):
    """The conditioning channels the GPU's UNet actually read: E2D mu (8..15) and distance features (16) of unet_in."""
    ui = p._session["unet_in"]
    c = ui.hi.view(torch.float16).float() if ui.f16 else ui.hi.float()
    if ui.lo is not None:
        c = c + ui.lo.float()
    c = c[:, 0, :, :, 8:17].permute(0, 3, 1, 2).contiguous().cpu()  # (N, 9, h, w)
    return c[:, :8], c[:, 8:9]


def _eps_errors(usd, rec, v_lat, feats, t_of_step):
    errs = []
    for step, x_t, eps_gpu, _ in rec:
        t = t_of_step(step)
        tb = torch.full((x_t.shape[0],), t, dtype=torch.long)
        eps_ref = ounet.unet_forward(usd, torch.cat([x_t, v_lat, feats], 1), tb)
        errs.append((step, t, rel_err(eps_gpu, eps_ref)))
    return errs


def _ddim50(case, ref, precision):
    usd, vsd, img, v2d, noise = case
    steps = 50
    pe = _predictor(usd, vsd, precision, 1000, graph=False)
    rec = EveryNth(5)
    out_eager = pe.predict_ddim(img.cuda(), v2d.cuda(), num_steps=steps, eta=0.0, noise=noise.cuda(), record=rec).cpu()
    ts = pe.ddim_timesteps(steps)
    assert len(rec) == 10
    v_lat, feats = _gpu_conditioning(pe)
    errs = _eps_errors(usd, rec, v_lat, feats, lambda i: ts[i])
    print(f"DDIM-50 {precision} per-step eps max-rel error on the GPU's own trajectory:", [(s_, t, f"{e:.2e}") for s_, t, e in errs])
    del pe
    pg = _predictor(usd, vsd, precision, 1000, graph=True)
    out = pg.predict_ddim(img.cuda(), v2d.cuda(), num_steps=steps, eta=0.0, noise=noise.cuda()).cpu()
    assert pg._session["graph"] is not None
    assert torch.equal(out, out_eager)                                     # the captured loop is the eager loop
    e = rel_l2(out, ref)
    print(f"DDIM-50 {precision} graph loop, B=2 11x256x256: final field rel-L2 vs the oracle trajectory = {e:.3e}")
    assert out.shape == ref.shape == (B, S, 3, SIZE, SIZE)
    assert (out[(img == 0).expand_as(out)] == 0).all()
    return max(e_ for _, _, e_ in errs), e


def test_ddim50_f16_graph_loop_vs_oracle(case, ddim50_oracle):
    """BASELINE configs[2]/[3] in the default 16-bit mode: both north-star bounds, with the measured margin held."""
    eps_err, field_err = _ddim50(case, ddim50_oracle, "f16")
    assert eps_err <= 2e-2 and field_err <= 1e-2, (eps_err, field_err)
    assert eps_err <= 5e-3 and field_err <= 5e-3, (eps_err, field_err)     # fp16 operands: ~8x inside the bf16 numbers


def test_ddim50_bf16_operands_measured(case, ddim50_oracle):
    """bf16 operands: the per-step bound holds, the 50-step field bound does NOT (1.8e-2 measured) -- recorded, not hidden;
    the assertion pins the measured level so a regression in either direction shows up."""
    eps_err, field_err = _ddim50(case, ddim50_oracle, "bf16")
    assert eps_err <= 2e-2, eps_err
    assert field_err <= 2.5e-2, field_err


def test_ddpm1000_fp32x_batch2_eps_every_50th_step(case):
    usd, vsd, img, v2d, noise = case
    T = 1000
    p = _predictor(usd, vsd, "fp32x", T, graph=False)
    rec = EveryNth(50)
    out = p.predict(img.cuda(), v2d.cuda(), noise=noise.cuda(), step_noise=StepNoise(7, T), record=rec).cpu()
    assert len(rec) == 20 and torch.isfinite(out).all()
    v_lat, feats = _gpu_conditioning(p)
    errs = _eps_errors(usd, rec, v_lat, feats, lambda i: T - 1 - i)
    print("DDPM-1000 fp32x per-step eps max-rel error on the GPU's own trajectory:", [(s_, t, f"{e:.2e}") for s_, t, e in errs])
    assert max(e for _, _, e in errs) <= 1e-3, errs
    # the conditioning itself (E2D + EDT + bilinear in the fp32-class mode) against the oracle's
    v_ref, f_ref = opred.conditioning(vsd, img, v2d, synth.NORM_FACTORS)
    assert rel_err(v_lat, v_ref) <= 1e-3 and rel_err(feats, f_ref) <= 1e-3


def test_ddpm100_fp32x_batch2_full_trajectory_vs_oracle(case):
    usd, vsd, img, v2d, noise = case
    T = 100
    zs = StepNoise(9, T)
    p = _predictor(usd, vsd, "fp32x", T, graph=False)
    out = p.predict(img.cuda(), v2d.cuda(), noise=noise.cuda(), step_noise=zs).cpu()
    ref = opred.predict(usd, vsd, img, v2d, noise, zs, norm_factors=synth.NORM_FACTORS, num_timesteps=T)
    e = rel_l2(out, ref)
    print(f"DDPM-100 fp32x, B=2 11x256x256, host-injected noise: final field rel-L2 vs the oracle trajectory = {e:.3e}")
    assert e <= 1e-2, e
