"""The committed golden vectors ARE what the unmodified reference computes: where the reference tree is present (the build
container; not the GPU box) both generator scripts are re-run into a scratch directory -- importing /root/reference,
seeded synthetic weights and inputs -- and every array is compared with the committed fixture for EQUALITY.  This pins
the fixtures (and through them the oracle and every GPU parity test) to the reference itself, not to a file nobody can
re-derive."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "Diffusion_model")), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("script, files", [
    ("make_golden.py", ("reference_keys.json", "scheduler.npz", "unet.npz", "vae.npz", "predict_ddim.npz", "predict_ddpm.npz")),
    ("make_train_golden.py", ("train_step.npz", "encode_target.npz", "train_from_fields.npz")),
])
def test_goldens_regenerate_bit_identically_from_the_reference(tmp_path, script, files):
    env = dict(os.environ, B2D_GOLDEN_OUT=str(tmp_path), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(GOLDEN, script)], env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert sorted(os.path.basename(p) for p in glob.glob(str(tmp_path / "*"))) == sorted(files)
    for name in files:
        new, old = str(tmp_path / name), os.path.join(GOLDEN, name)
        if name.endswith(".json"):
            assert json.load(open(new)) == json.load(open(old)), name
            continue
        a, b = np.load(new), np.load(old)
        assert sorted(a.files) == sorted(b.files), name
        for k in a.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, (name, k)
            assert np.array_equal(a[k], b[k]), (name, k)
