"""Host logic of diffusion_model_project_b200.checkpoint (CPU): the reference's checkpoint-directory layout
(Predictor.from_directory predictor.py:222-250, dual-VAE directories :476-566, vae_log.json norm factors :363-369)
and the state-dict hand-over from a constructed reference predictor (INTEGRATION.md `use_b200()`)."""
import json
import os
import sys
import types

import pytest
import torch

from diffusion_model_project_b200 import checkpoint, synth

REF = "/root/reference"
SMALL = dict(synth.UNET_KWARGS, features=[64, 128], attention="2..2")


def _write_dirs(tmp_path, unet_kwargs, with_saved_vae=True):
    usd = synth.synth_unet_state(seed=0, **unet_kwargs)
    vsd = synth.synth_vae_state(seed=1, branches=("encoder_2d", "encoder_3d", "decoder_3d"))
    enc, dec, run = tmp_path / "stage2_e2d", tmp_path / "stage1_3d", tmp_path / "run"
    for d in (enc, dec, run):
        d.mkdir()
    # stage 2 holds E2D under `encoder_2d.`; stage 1 holds the 3D branch under the standard-VAE names `encoder.` / `decoder.`
    torch.save({k: v for k, v in vsd.items() if k.startswith("encoder_2d.")}, enc / "best_model.pt")
    stage1 = {k.replace("encoder_3d.", "encoder.").replace("decoder_3d.", "decoder."): v for k, v in vsd.items()
              if k.startswith(("encoder_3d.", "decoder_3d."))}
    torch.save(stage1, dec / "vae.pt")
    (dec / "vae_log.json").write_text(json.dumps({"norm_factors": synth.NORM_FACTORS, "latent_channels": 8, "in_channels": 3}))
    kw = dict(model_name="UNet", model_kwargs={k: v for k, v in unet_kwargs.items() if k != "time_embedding_dim"},
              distance_transform=True, vae_encoder_path=str(enc), vae_decoder_path=str(dec), num_slices=11, num_timesteps=1000)
    (run / "log.json").write_text(json.dumps({"params": {"training": {"predictor_type": "latent-diffusion", "predictor": kw}}}))
    state = {"model." + k: v for k, v in usd.items()}
    state["distance_transform"] = torch.tensor([1.0])
    state["scheduler.betas"] = torch.linspace(1e-4, 0.02, 1000)
    if with_saved_vae:
        state.update({"vae." + k: v for k, v in vsd.items()})
        state["normalizer.input.scale_factors"] = torch.tensor([1.0])
        state["normalizer.output.scale_factors"] = torch.tensor(synth.NORM_FACTORS)
    torch.save(state, run / "model.pt")
    return run, usd, vsd


@pytest.mark.parametrize("with_saved_vae", [True, False])
def test_load_directory_reads_the_reference_layout(tmp_path, with_saved_vae):
    run, usd, vsd = _write_dirs(tmp_path, SMALL, with_saved_vae)
    spec = checkpoint.load_directory(str(run))
    assert spec.model_kwargs["time_embedding_dim"] == 64 and spec.model_kwargs["features"] == [64, 128]
    assert spec.num_slices == 11 and spec.num_timesteps == 1000 and spec.distance_transform is True
    assert spec.norm_factors == pytest.approx(synth.NORM_FACTORS)
    assert set(spec.unet_state) == set(usd) and all(torch.equal(spec.unet_state[k], usd[k]) for k in usd)
    assert set(spec.vae_state) == set(vsd) and all(torch.equal(spec.vae_state[k], vsd[k]) for k in vsd)


def test_load_directory_errors(tmp_path):
    run, _, _ = _write_dirs(tmp_path, SMALL)
    log = json.loads((run / "log.json").read_text())
    log["params"]["training"]["predictor_type"] = "deterministic"
    (run / "log.json").write_text(json.dumps(log))
    with pytest.raises(ValueError, match="Unknown or unsupported predictor type"):
        checkpoint.load_directory(str(run))
    log["params"]["training"]["predictor_type"] = "latent-diffusion"
    log["params"]["training"]["predictor"]["vae_decoder_path"] = str(tmp_path / "missing")
    (run / "log.json").write_text(json.dumps(log))
    with pytest.raises(FileNotFoundError):
        checkpoint.load_directory(str(run))
    del log["params"]["training"]["predictor"]["vae_decoder_path"]
    (run / "log.json").write_text(json.dumps(log))
    with pytest.raises(ValueError, match="VAE path must be provided"):
        checkpoint.load_directory(str(run))


def test_spec_from_a_reference_shaped_object():
    """Duck-typed stand-in with exactly the attributes the reference predictor keeps (predictor.py:129-148, 301-340)."""
    usd = synth.synth_unet_state(seed=0, **SMALL)
    vsd = synth.synth_vae_state(seed=1)
    model = types.SimpleNamespace(in_channels=17, out_channels=8, features=[64, 128], kernel_size=3, padding_mode="zeros",
                                  _activation="silu", _final_activation=None, attention="2..2", dropout=0.0, time_embedding_dim=64,
                                  state_dict=lambda: usd)
    pred = types.SimpleNamespace(model=model, vae=types.SimpleNamespace(state_dict=lambda: vsd), vae_is_dual=True,
                                 distance_transform=torch.nn.Parameter(torch.tensor([1.0]), requires_grad=False),
                                 normalizer={"output": types.SimpleNamespace(scale_factors=torch.tensor(synth.NORM_FACTORS))},
                                 num_slices=11, num_timesteps=1000)
    spec = checkpoint.spec_from_reference(pred)
    assert spec.model_kwargs == {k: SMALL[k] for k in spec.model_kwargs}
    assert set(spec.unet_state) == set(usd) and set(spec.vae_state) == set(vsd)
    assert spec.norm_factors == pytest.approx(synth.NORM_FACTORS) and spec.distance_transform is True
    pred.vae_is_dual = False
    with pytest.raises(NotImplementedError):
        checkpoint.spec_from_reference(pred)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_spec_from_the_real_reference_predictor(tmp_path):
    """The unmodified reference, constructed from the same directories through its own __init__ / load_weights: the
    spec taken from its modules carries the reference's own key list (tests/golden/reference_keys.json) and the same
    tensors `load_directory` reads from disk."""
    run, usd, vsd = _write_dirs(tmp_path, SMALL)
    saved = list(sys.path)
    sys.path[:0] = [REF, os.path.join(REF, "Diffusion_model")]
    sys.dont_write_bytecode = True
    try:
        from src.predictor import LatentDiffusionPredictor
        ref = LatentDiffusionPredictor.from_directory(str(run), device="cpu")
    finally:
        sys.path[:] = saved
    a = checkpoint.spec_from_reference(ref)
    b = checkpoint.load_directory(str(run))
    assert a.model_kwargs == b.model_kwargs and a.num_slices == b.num_slices and a.num_timesteps == b.num_timesteps
    assert a.norm_factors == pytest.approx(b.norm_factors) and a.distance_transform == b.distance_transform
    assert set(a.unet_state) == set(b.unet_state) and all(torch.equal(a.unet_state[k], b.unet_state[k]) for k in b.unet_state)
    for k, v in b.vae_state.items():  # the reference's DualBranchVAE also owns decoder_2d; the sampling path never reads it
        assert torch.equal(a.vae_state[k], v)
