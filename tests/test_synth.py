"""synth.py enumerates the reference's state-dict keys/shapes without importing it; the list
captured from the real modules (tests/golden/reference_keys.json) pins that."""
import json
import os

import torch

from diffusion_model_project_b200 import synth


def test_unet_keys_match_reference(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "reference_keys.json")))["unet"]
    spec = synth.unet_spec(**synth.UNET_KWARGS)
    assert [k for k, _ in ref] == list(spec.keys())
    assert [tuple(s) for _, s in ref] == list(spec.values())
    assert sum(torch.Size(s).numel() for s in spec.values()) == 139_810_952


def test_vae_keys_match_reference(golden_dir):
    ref = dict((k, tuple(s)) for k, s in json.load(open(os.path.join(golden_dir, "reference_keys.json")))["vae"])
    spec = synth.dual_vae_spec()
    assert list(spec.keys()) == list(ref.keys())
    assert all(spec[k] == ref[k] for k in spec)
    enc = synth.encoder_spec()
    dec = synth.decoder_spec()
    assert sum(torch.Size(s).numel() for s in enc.values()) == 35_354_768
    assert sum(torch.Size(s).numel() for s in dec.values()) == 41_715_459


def test_synth_deterministic_and_nonzero():
    a = synth.synth_state_dict(synth.decoder_spec(), seed=3)
    b = synth.synth_state_dict(synth.decoder_spec(), seed=3)
    for k in a:
        assert torch.equal(a[k], b[k])
    u = synth.synth_unet_state(seed=0, features=[64, 128], attention="2..2")
    assert u["final_conv.weight"].abs().max() > 0
    assert u["encoder.1.1.proj_out.weight"].abs().max() > 0


def test_attention_expr():
    assert synth.attention_heads("3..2", 5) == [None, None, 2, 2, 2]
    assert synth.attention_heads("", 4) == [None] * 4
    assert synth.attention_heads("1.1.1", 3) == [1, None, None]
