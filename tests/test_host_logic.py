"""Host-side logic that needs no GPU: weight repacking / tap tables (checked through a CPU emulation
of the engine's GEMM view against torch convs), coefficient rows, time-embedding table, ABI surface."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

from diffusion_model_project_b200 import _lib, engine, synth
from diffusion_model_project_b200.scheduler import B200Scheduler
from diffusion_model_project_b200.unet import B200UNet
from emu import emulate_conv
from oracle import unet as ounet
from oracle.scheduler import OracleScheduler

torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cl(x):  # planar (N,C,D,H,W) -> channels-last padded fp32
    N, C, D, H, W = x.shape
    out = torch.zeros(N, D, H, W, engine.pad64(C))
    out[..., :C] = x.permute(0, 2, 3, 4, 1)
    return out


def test_pack_conv2d_and_concat():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 24, 1, 6, 6, generator=g)
    w = torch.randn(10, 24, 3, 3, generator=g)
    pw = engine.pack_conv2d(w, [24], None, "cpu")
    pw.w = pw.w.float()
    pw.w = engine._pack_taps(w.permute(0, 2, 3, 1).reshape(10, 9, 24), [24])[0]
    got = emulate_conv([_cl(x)], pw, 10)
    ref = F.conv2d(x[:, :, 0], w, padding=1)[:, :, None]
    assert (got - ref).abs().max() < 1e-4
    # torch.cat((skip, up), 1) as two K segments
    pw2 = engine.pack_conv2d(w, [8, 16], None, "cpu")
    pw2.w = engine._pack_taps(w.permute(0, 2, 3, 1).reshape(10, 9, 24), [8, 16])[0]
    got2 = emulate_conv([_cl(x[:, :8]), _cl(x[:, 8:])], pw2, 10)
    assert (got2 - ref).abs().max() < 1e-4
    assert pw2.kbase == [0, 9 * 64] and pw2.cin_pad == [64, 64]


def test_pack_conv3d_and_down():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 5, 3, 8, 8, generator=g)
    w = torch.randn(7, 5, 3, 3, 3, generator=g)
    b = torch.randn(7, generator=g)
    for down in (False, True):
        pw = engine.pack_conv3d(w, b, "cpu", down=down)
        pw.w = engine._pack_taps(w.permute(0, 2, 3, 4, 1).reshape(7, 27, 5), [5])[0]
        got = emulate_conv([_cl(x)], pw, 7, stride=2 if down else 1)
        if down:
            ref = F.conv3d(F.pad(x, (0, 1, 0, 1, 1, 1)), w, b, stride=(1, 2, 2))  # encoder.py:76-81
        else:
            ref = F.conv3d(x, w, b, padding=1)
        assert (got - ref).abs().max() < 1e-4


def test_pack_convT_and_linear_fold():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 12, 1, 4, 4, generator=g)
    w = torch.randn(12, 6, 2, 2, generator=g)
    b = torch.randn(6, generator=g)
    pw = engine.pack_convT2x2(w, b, "cpu")
    pw.w = engine._pack_taps(w.permute(2, 3, 1, 0).reshape(24, 1, 12), [12])[0]
    got = emulate_conv([_cl(x)], pw, 6, nphase=4)
    ref = F.conv_transpose2d(x[:, :, 0], w, b, stride=2)[:, :, None]
    assert (got - ref).abs().max() < 1e-4


def test_split_hi_lo():
    w = torch.randn(1000) * 3
    hi, lo = engine.split_hi_lo(w)
    assert ((hi.float() + lo.float()) - w).abs().max() <= 2 ** -15 * w.abs().max()


def test_scheduler_coef_rows_match_oracle():
    s = B200Scheduler(1000, device="cpu")
    o = OracleScheduler(1000)
    for k in s._names:
        assert torch.equal(getattr(s, k), getattr(o, k)), k
    g = torch.Generator().manual_seed(3)
    x, e, z = (torch.randn(64, generator=g) for _ in range(3))

    def apply(row, kind, x, e, z, lo, hi):
        a, b, c1, c2, sg = row[:5]
        x0 = torch.clamp((x - b * e) / a, lo, hi)
        out = c1 * x0 + c2 * (x if kind == 0 else e)
        return out + sg * z if sg != 0 else out

    rows = s.ddpm_coef_rows([999, 500, 1, 0])
    for r, t in zip(rows, [999, 500, 1, 0]):
        assert torch.equal(apply(r, 0, x, e, z, -30.0, 30.0), o.p_sample(e, x, t, z, True, (-30.0, 30.0)))
    ts = [999, 978, 20, 0]
    for eta in (0.0, 0.6):
        rows = s.ddim_coef_rows(ts, eta)
        for i, t in enumerate(ts):
            tp = ts[i + 1] if i + 1 < len(ts) else -1
            assert torch.equal(apply(rows[i], 1, x, e, z, -30.0, 30.0), o.ddim_sample(e, x, t, tp, eta, (-30.0, 30.0), noise=z))


def test_time_table_matches_oracle():
    kw = dict(synth.UNET_KWARGS, features=[64, 128], attention="2..2")
    sd = synth.synth_unet_state(seed=5, **kw)
    m = B200UNet(**kw, device="cpu", num_timesteps=50)
    m._build_time_table({k: v.float() for k, v in sd.items()})
    t = torch.tensor([0, 7, 49])
    temb = ounet.time_embedding(sd, t, 64)
    for p, col in m._temb_cols.items():
        ref = F.linear(F.silu(temb), sd[f"{p}.time_mlp.1.weight"], sd[f"{p}.time_mlp.1.bias"])
        got = m.temb_table[t, col:col + ref.shape[1]]
        assert (got - ref).abs().max() < 1e-5
    assert m.temb_table.shape == (50, 64 + 128 + 256 + 128 + 64)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b2d.h")).read()
    declared = set(re.findall(r"B2D_API\s+[\w\s\*]+?\b(b2d_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.b2d_version() == _lib.ABI_VERSION == 8


def test_abi_rejects_bad_arguments_without_a_gpu():
    lib = _lib.lib()
    assert lib.b2d_scheduler_step(0, None, None, None, None, 16, None, None, 0, 0, 0, 0.0, 0.0, None, 0, 0, 0, None, None, 0, None) == -1
    assert b"null" in lib.b2d_last_error()
    assert lib.b2d_attention(1, None, 1, None, 1, 16, 100, 2, 0, None) == -1  # d_head 50 not a multiple of 64
    d = _lib.ConvDesc()
    h = ctypes.c_void_p()
    assert lib.b2d_conv_plan_create(ctypes.byref(d), ctypes.byref(h)) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, "x")


def test_product_path_refuses_cpu_tensors():
    kw = dict(synth.UNET_KWARGS, features=[64, 128], attention="")
    m = B200UNet(**kw, device="cpu")
    with pytest.raises(ValueError):
        m.forward(torch.zeros(1, 17, 32, 32), None)  # models.py:139-140
    m._w = {"x": 1}
    with pytest.raises(RuntimeError):
        m.forward(torch.zeros(1, 17, 32, 32), torch.zeros(1, dtype=torch.long))
    s = B200Scheduler(10, device="cpu")
    with pytest.raises(RuntimeError):
        s.p_sample(torch.zeros(4), torch.zeros(4), 3)
    with pytest.raises(NotImplementedError):
        B200UNet(padding_mode="reflect")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.B2DError):
        _lib.lib()


def test_upsample_folded_conv_weights_match_reference_math():
    """decoder.py:46-47: Conv3d(Upsample(scale=(1,2,2))(x)) == four phase convs with 2x2x3 taps on the low-res map
    (engine.pack_conv3d_upsampled).  Checked on the CPU with the packed matrices themselves."""
    import torch.nn.functional as F
    from diffusion_model_project_b200 import engine
    g = torch.Generator().manual_seed(3)
    ci, co = 64, 64
    w = torch.randn(co, ci, 3, 3, 3, generator=g) * 0.05
    b = torch.randn(co, generator=g)
    x = torch.randn(1, ci, 3, 6, 5, generator=g)
    ref = F.conv3d(F.interpolate(x, scale_factor=(1, 2, 2), mode="nearest"), w, b, padding=1)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        pw = engine.pack_conv3d_upsampled(w, b, "cpu", py, px)
        assert len(pw.taps) == 12 and pw.ktot == 12 * 64
        wk = torch.zeros(co, ci, 3, 3, 3)
        mat = pw.w.float()[:co].reshape(co, 12, 64)
        for t, (dz, dy, dx) in enumerate(pw.taps):
            wk[:, :, dz + 1, dy + 1, dx + 1] = mat[:, t, :ci]
        got = F.conv3d(x, wk, b, padding=1)
        want = ref[:, :, :, py::2, px::2]
        assert (got - want).abs().max() <= 2e-2 * want.abs().max()  # bf16-rounded packed weights


def test_zfold_conv_out_weights_match_reference_math():
    """decoder.py:71: Conv3d(C -> 3, k3, pad 1) == a per-slice 3x3 conv with rows (kz, co) followed by a gather over z
    (engine.pack_conv3d_zfold + b2d_zfold_combine).  Checked on the CPU with the packed matrix itself."""
    import torch.nn.functional as F
    from diffusion_model_project_b200 import engine
    g = torch.Generator().manual_seed(5)
    ci, co, D = 64, 3, 4
    w = torch.randn(co, ci, 3, 3, 3, generator=g) * 0.05
    b = torch.randn(co, generator=g)
    x = torch.randn(1, ci, D, 6, 5, generator=g)
    ref = F.conv3d(x, w, b, padding=1)
    pw = engine.pack_conv3d_zfold(w, "cpu")
    assert len(pw.taps) == 9 and all(dz == 0 for dz, _, _ in pw.taps) and pw.ktot == 9 * 64 and pw.bias is None
    mat = pw.w.float()[:12].reshape(12, 9, 64)
    w2d = torch.zeros(12, ci, 3, 3)
    for t, (_, dy, dx) in enumerate(pw.taps):
        w2d[:, :, dy + 1, dx + 1] = mat[:, t, :ci]
    P = torch.stack([F.conv2d(x[:, :, z], w2d, padding=1) for z in range(D)], dim=2)  # [1, 12, D, H, W]
    got = torch.zeros_like(ref)
    for z in range(D):
        for kz in range(3):
            zi = z + kz - 1
            if 0 <= zi < D:
                got[:, :, z] += P[:, kz * 4:kz * 4 + co, zi]
    got += b.view(1, co, 1, 1, 1)
    assert (got - ref).abs().max() <= 2e-2 * ref.abs().max()  # bf16-rounded packed weights


def test_zstack_conv_in_weights_match_reference_math():
    """encoder.py:30 / decoder.py:31: Conv3d(C <= 21 -> Cout, k3, pad 1) == a 3x3 in-plane conv over the z-stacked input
    (channel kz*C + c = slice z+kz-1; engine.pack_conv3d_zstack + b2d_zstack_cl).  CPU check with the packed matrix."""
    import torch.nn.functional as F
    from diffusion_model_project_b200 import engine
    g = torch.Generator().manual_seed(7)
    ci, co, D = 3, 16, 4
    w = torch.randn(co, ci, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(co, generator=g)
    x = torch.randn(1, ci, D, 6, 5, generator=g)
    ref = F.conv3d(x, w, b, padding=1)
    pw = engine.pack_conv3d_zstack(w, b, "cpu")
    assert len(pw.taps) == 9 and pw.ktot == 9 * 64 and pw.cin_pad == [64]
    mat = pw.w.float()[:co].reshape(co, 9, 64)
    w2d = torch.zeros(co, 3 * ci, 3, 3)
    for t, (_, dy, dx) in enumerate(pw.taps):
        w2d[:, :, dy + 1, dx + 1] = mat[:, t, :3 * ci]
    xp = F.pad(x, (0, 0, 0, 0, 1, 1))                                                 # zero slices at z = -1 and z = D
    xs = torch.cat([xp[:, :, kz:kz + D] for kz in range(3)], dim=1)                     # [1, 3*ci, D, H, W], channel kz*ci + c
    got = torch.stack([F.conv2d(xs[:, :, z], w2d, b, padding=1) for z in range(D)], dim=2)
    assert (got - ref).abs().max() <= 2e-2 * ref.abs().max()  # bf16-rounded packed weights


def test_only_the_checkers_import_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs (one function) may import
    it -- never the product package or the measurement tools."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    offenders = []
    for sub in ("diffusion_model_project_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith(".py") and pat.search(open(os.path.join(dirpath, f)).read()):
                    offenders.append(os.path.join(sub, f))
    assert offenders == []
    bench = open(os.path.join(root, "bench.py")).read()
    hits = [m.start() for m in pat.finditer(bench)]
    fn = bench.index("def cpu_reference_prediction"), bench.index("def run_reference")
    assert len(hits) == 1 and fn[0] < hits[0] < fn[1]          # inside cpu_reference_prediction() only
    entry = open(os.path.join(root, "__graft_entry__.py")).read()
    assert all(entry.rfind("\ndef ", 0, m.start()) >= entry.index("\ndef build") for m in pat.finditer(entry))  # inside build() / smoke()
