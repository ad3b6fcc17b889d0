"""Pin the CPU oracle against vectors produced by the unmodified reference
(tests/golden/make_golden.py) and the known-answer values of SURVEY.md section 8(c)."""
import os
import sys

import numpy as np
import pytest
import torch

from diffusion_model_project_b200 import synth
from oracle import predictor as opred
from oracle import unet as ounet
from oracle import vae as ovae
from oracle.scheduler import OracleScheduler, ddim_timesteps

torch.set_grad_enabled(False)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_scheduler_known_answers():
    s = OracleScheduler(1000)
    idx = [0, 1, 500, 999]
    np.testing.assert_allclose(s.betas[idx].numpy(), [1.0e-4, 1.1991992e-4, 1.0059960e-2, 2.0e-2], rtol=2e-7)
    np.testing.assert_allclose(s.alphas_cumprod[idx].numpy(), [0.99989998, 0.99978012, 0.07779666, 4.0358296e-5], rtol=2e-7)
    np.testing.assert_allclose(s.sqrt_alphas_cumprod[999].item(), 0.0063528181, rtol=2e-7)
    np.testing.assert_allclose(s.posterior_variance[idx].numpy(), [1e-20, 5.4531876e-5, 1.0051336e-2, 1.9999983e-2], rtol=2e-7)
    np.testing.assert_allclose(s.posterior_mean_coef1[idx].numpy(), [1.0, 0.54529148, 3.0580570e-3, 1.2835149e-4], rtol=2e-7)
    np.testing.assert_allclose(s.posterior_mean_coef2[idx].numpy(), [0.0, 0.45470849, 0.99410433, 0.98994869], rtol=2e-7)
    ts = ddim_timesteps(1000, 50)
    assert ts[:5] == [999, 978, 958, 937, 917] and ts[-5:] == [81, 61, 40, 20, 0] and len(ts) == 50


def test_scheduler_vs_reference(golden_dir):
    g = _load(golden_dir, "scheduler.npz")
    s = OracleScheduler(1000)
    for k in ("betas", "alphas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
              "sqrt_one_minus_alphas_cumprod", "posterior_variance", "posterior_log_variance",
              "posterior_mean_coef1", "posterior_mean_coef2"):
        assert np.array_equal(getattr(s, k).numpy(), g[k]), k
    x, eps, z = (torch.from_numpy(g[k]) for k in ("x", "eps", "z"))
    for t in g["p_ts"].tolist():
        assert np.array_equal(s.p_sample(eps, x, t, z, True, (-30.0, 30.0)).numpy(), g[f"p_sample_{t}"])
        assert np.array_equal(s.q_sample(x, t, eps).numpy(), g[f"q_sample_{t}"])
        assert np.array_equal(s.predict_x0_from_noise(x, t, eps).numpy(), g[f"x0_{t}"])
    assert np.array_equal(s.p_sample(eps, x, 999, z).numpy(), g["p_sample_default_clip_999"])
    for t, tp in g["ddim_pairs"].tolist():
        assert np.array_equal(s.ddim_sample(eps, x, t, tp, 0.0).numpy(), g[f"ddim_{t}_{tp}"])
        assert np.array_equal(s.ddim_sample(eps, x, t, tp, 0.5, noise=z).numpy(), g[f"ddim_eta05_{t}_{tp}"])


def test_unet_vs_reference(golden_dir):
    g = _load(golden_dir, "unet.npz")
    sd = synth.synth_unet_state(seed=0)
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(2, 17, 32, 32, generator=gen)
    eps = ounet.unet_forward(sd, x, torch.from_numpy(g["t"]))
    ref = torch.from_numpy(g["eps"])
    assert (eps - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def test_vae_vs_reference(golden_dir):
    g = _load(golden_dir, "vae.npz")
    sd = synth.synth_vae_state(seed=1)
    gen = torch.Generator().manual_seed(13)
    xv = torch.randn(1, 3, 3, 32, 32, generator=gen)
    z, (mu, logvar) = ovae.encode_2d_deterministic(sd, xv)
    zl = torch.randn(1, 8, 3, 8, 8, generator=gen)
    dec = ovae.decode_3d(sd, zl)
    for a, b in ((mu, g["mu"]), (logvar, g["logvar"]), (dec, g["dec"])):
        b = torch.from_numpy(b)
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item()
    assert z is mu


@pytest.mark.parametrize("eta", [0.0, 0.7])
def test_predict_ddim_vs_reference(golden_dir, eta):
    g = _load(golden_dir, "predict_ddim.npz")
    usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
    img, v2d = synth.synth_inputs(1, num_slices=2, size=128, seed=2024)
    noise = synth.synth_noise(1, num_slices=2, latent_size=32, seed=42)
    rec = []
    zs = None
    if eta > 0:
        gen = torch.Generator().manual_seed(99)
        zs = [torch.randn(2, 8, 32, 32, generator=gen) for _ in range(3)]
    out = opred.predict_ddim(usd, vsd, img, v2d, noise, num_steps=3, eta=eta, norm_factors=synth.NORM_FACTORS,
                             record=rec, step_noise=zs)
    ref = torch.from_numpy(g["out"] if eta == 0.0 else g["out_eta07"])
    assert out.shape == ref.shape == (1, 2, 3, 128, 128)
    assert ((out - ref).norm() / ref.norm()).item() <= 1e-4
    if eta == 0.0:
        eps_ref = torch.from_numpy(g["eps_steps"])
        for i, (_, _, eps, _) in enumerate(rec):
            assert (eps - eps_ref[i]).abs().max().item() <= 1e-4 * eps_ref[i].abs().max().item()


def test_predict_ddpm_vs_reference(golden_dir):
    g = _load(golden_dir, "predict_ddpm.npz")
    usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
    img, v2d = synth.synth_inputs(1, num_slices=2, size=128, seed=2024)
    noise = synth.synth_noise(1, num_slices=2, latent_size=32, seed=42)
    gen = torch.Generator().manual_seed(100)
    zs = [torch.randn(2, 8, 32, 32, generator=gen) for _ in range(12)]
    out = opred.predict(usd, vsd, img, v2d, noise, zs, norm_factors=synth.NORM_FACTORS, num_timesteps=12)
    ref = torch.from_numpy(g["out"])
    assert ((out - ref).norm() / ref.norm()).item() <= 1e-4


def test_training_step_vs_reference(golden_dir):
    """SURVEY.md section 8 row f4 (next row): loss, noise prediction, parameter gradients and the Adam update of ONE UNet
    training step, oracle (autograd through the functional oracle UNet + restated loss / Adam) against the unmodified
    reference modules (tests/golden/make_train_golden.py)."""
    import importlib.util
    from oracle import train as otrain
    spec = importlib.util.spec_from_file_location("make_train_golden", os.path.join(golden_dir, "make_train_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    saved = list(sys.path)
    try:
        spec.loader.exec_module(mod)  # only train_inputs() / FULL are used; the reference is not imported at module level
    finally:
        sys.path[:] = saved
    g = _load(golden_dir, "train_step.npz")
    sd = synth.synth_unet_state(seed=0)
    x_start, cond, feats, noise, t = mod.train_inputs()
    assert np.array_equal(t.numpy(), g["t"])
    loss, grads, pred = otrain.training_loss_and_grads(sd, x_start, cond, feats, t, noise)
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    ref_pred = torch.from_numpy(g["pred"])
    assert (pred - ref_pred).abs().max().item() <= 2e-5 * ref_pred.abs().max().item()
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(sd.keys())
    for n, ref_norm in zip(names, g["grad_norms"]):
        got = float(grads[n].double().norm())
        assert abs(got - ref_norm) <= 2e-3 * max(ref_norm, 1e-6), (n, got, ref_norm)
    for k in mod.FULL:
        ref_g = torch.from_numpy(g[f"grad::{k}"])
        assert (grads[k] - ref_g).abs().max().item() <= 2e-3 * max(ref_g.abs().max().item(), 1e-12), k
        # the Adam restatement in isolation: reference gradient in, reference parameter delta out
        p0 = sd[k]
        p1, m, v = otrain.adam_step(p0, ref_g, torch.zeros_like(p0), torch.zeros_like(p0), step=1, lr=1e-4)
        ref_d = torch.from_numpy(g[f"delta::{k}"])
        assert ((p1 - p0) - ref_d).abs().max().item() <= 2e-8, k
    # the loss restatement on its own, 5-D form and channel weights (metrics.py:362-396)
    a = torch.randn(2, 3, 4, 5, 6, generator=torch.Generator().manual_seed(1))
    b = torch.randn(2, 3, 4, 5, 6, generator=torch.Generator().manual_seed(2))
    w = torch.tensor([1.0, 2.0, 0.5])
    mse = ((a - b) ** 2).mean(dim=(-3, -2, -1)) / ((b ** 2).mean(dim=(-3, -2, -1)) + 1e-8)
    want = (mse * w[None] / w.sum()).mean(dim=-1).mean()
    assert torch.allclose(otrain.normalized_mse_loss_per_component(a, b, weight_per_channel=w), want, rtol=1e-6, atol=0)
    with pytest.raises(ValueError):
        otrain.normalized_mse_loss_per_component(a[0, 0], b[0, 0])


def _load_train_golden_module(golden_dir):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_train_golden", os.path.join(golden_dir, "make_train_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    saved = list(sys.path)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved
    return mod


def test_encode_target_vs_reference(golden_dir):
    """LatentDiffusionPredictor.encode_target (predictor.py:1042-1085) -- oracle restatement
    against the unmodified reference with the seeded E3D weights (tests/golden/make_train_golden.py)."""
    mod = _load_train_golden_module(golden_dir)
    g = _load(golden_dir, "encode_target.npz")
    vsd = synth.synth_vae_state(seed=1, branches=("encoder_2d", "encoder_3d", "decoder_3d"))
    lat = opred.encode_target(vsd, mod.target_inputs(), synth.NORM_FACTORS)
    ref = torch.from_numpy(g["latents"])
    assert lat.shape == ref.shape == (2, 3, 8, 8, 8)
    assert (lat - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def test_training_step_from_fields_vs_reference(golden_dir):
    """The caller of row f4: the reference's training loop body from the FIELDS (helper.py:277-430: encode_target, then
    predictor.forward = frozen-VAE conditioning + randint t + q_sample + UNet, predictor.py:636-751; criterion; backward)
    -- oracle composition against the unmodified reference predictor (tests/golden/train_from_fields.npz)."""
    from oracle import train as otrain
    mod = _load_train_golden_module(golden_dir)
    g = _load(golden_dir, "train_from_fields.npz")
    img, v2d, target, noise, seed = mod.field_inputs()
    torch.manual_seed(seed)
    t = torch.randint(0, 1000, (2,)).long()
    assert np.array_equal(t.numpy(), g["t"])
    usd = synth.synth_unet_state(seed=0)
    vsd = synth.synth_vae_state(seed=1, branches=("encoder_2d", "encoder_3d", "decoder_3d"))
    loss, grads, pred, (x_start, _, _) = otrain.training_step_from_fields(usd, vsd, img, v2d, target, t, noise, synth.NORM_FACTORS)
    ref_lat = torch.from_numpy(g["latents"])
    assert (x_start.reshape(ref_lat.shape) - ref_lat).abs().max().item() <= 2e-5 * ref_lat.abs().max().item()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    ref_pred = torch.from_numpy(g["pred"])
    assert (pred - ref_pred).abs().max().item() <= 5e-5 * ref_pred.abs().max().item()
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(usd.keys())
    for n, ref_norm in zip(names, g["grad_norms"]):
        got = float(grads[n].double().norm())
        assert abs(got - ref_norm) <= 2e-3 * max(ref_norm, 1e-6), (n, got, ref_norm)
    for k in ("final_conv.weight", "encoder.0.0.block1.conv.weight", "bottleneck.block2.norm.weight"):
        ref_g = torch.from_numpy(g[f"grad::{k}"])
        assert (grads[k] - ref_g).abs().max().item() <= 2e-3 * max(ref_g.abs().max().item(), 1e-12), k
