"""Memory-bound kernels, attention core, EDT and bilinear resize vs plain PyTorch / SciPy."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffusion_model_project_b200 import _lib, engine
from diffusion_model_project_b200.engine import new_act
from util import bf16_round, fmt_round, from_act, no_tf32, rel_err, stats_ref, to_act

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"


def _s():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("shape,groups", [((3, 64, 1, 8, 8), 1), ((2, 128, 3, 8, 8), 32), ((11, 2048, 1, 2, 2), 1), ((1, 512, 2, 4, 4), 32)])
def test_gn_apply_silu_temb(shape, groups, f16):
    """f16: input and output stored as IEEE fp16 (the default 16-bit mode); the tolerance follows the storage format."""
    g = torch.Generator().manual_seed(sum(shape))
    N, C = shape[:2]
    tol = 1.5e-3 if f16 else 1e-2
    x = fmt_round((torch.randn(*shape, generator=g) * 2 + 0.5).to(DEV), f16)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(C, generator=g)).to(DEV)
    table = torch.randn(5, C + 7, generator=g).to(DEV)
    rows = torch.randint(0, 5, (N,), generator=g).to(torch.int32).to(DEV)
    st = stats_ref(x, groups).to(DEV).contiguous()
    xa = to_act(x, f16=f16)
    ya = new_act(*xa.shape, DEV, f16=f16)
    st_out = torch.zeros(N, 2, dtype=torch.float64, device=DEV)
    engine.gn_apply(xa, ya, st, C // groups, gamma, beta, True, _s(), temb=table, temb_row=rows, temb_row_stride=1, temb_col=7,
                    stats_out=st_out)
    ref = F.silu(F.group_norm(x, groups, gamma, beta, eps=1e-5))
    ref = ref + table[rows.long(), 7:][:, :, None, None, None]
    assert rel_err(from_act(ya, C), ref) < tol
    sref = stats_ref(ref, 1)[:, 0]
    assert ((st_out - sref).abs().max() / sref.abs().max()).item() < 5e-3
    # no activation / no temb variant (attention pre-norm)
    engine.gn_apply(xa, ya, st, C // groups, gamma, beta, False, _s())
    assert rel_err(from_act(ya, C), F.group_norm(x, groups, gamma, beta, eps=1e-5)) < tol


@pytest.mark.parametrize("f16", [False, True])
def test_maxpool_and_upsample(f16):
    g = torch.Generator().manual_seed(1)
    x = fmt_round(torch.randn(3, 128, 1, 8, 12, generator=g).to(DEV), f16)
    xa = to_act(x, f16=f16)
    ya = new_act(3, 1, 4, 6, 128, DEV, f16=f16)
    st = torch.zeros(3, 2, dtype=torch.float64, device=DEV)
    engine.maxpool_stats(xa, ya, st, _s())
    ref = F.max_pool2d(x[:, :, 0], 2, 2)
    assert torch.equal(from_act(ya, 128)[:, :, 0], ref)
    sref = stats_ref(ref[:, :, None], 1)[:, 0]
    assert ((st - sref).abs().max() / sref.abs().max()).item() < 1e-5
    x3 = fmt_round(torch.randn(2, 64, 3, 4, 4, generator=g).to(DEV), f16)
    ua = new_act(2, 3, 8, 8, 64, DEV, f16=f16)
    engine.upsample2x(to_act(x3, f16=f16), ua, _s())
    assert torch.equal(from_act(ua, 64), F.interpolate(x3, scale_factor=(1, 2, 2)))  # decoder.py:46


def test_layout_conversions_roundtrip():
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(4, 3, 100, generator=g) * 0.01).to(DEV)
    scale = torch.tensor([0.01, 0.005, 0.002], device=DEV)
    y = torch.zeros(4, 100, 64, dtype=torch.bfloat16, device=DEV)
    _lib.call("b2d_planar_to_cl", x.data_ptr(), y.data_ptr(), None, 4, 3, 100, 64, 8, scale.data_ptr(), 0, _s())
    ref = (x / scale.view(1, 3, 1)).to(torch.bfloat16)
    assert torch.equal(y[:, :, 8:11], ref.permute(0, 2, 1))
    assert y[:, :, :8].abs().max() == 0 and y[:, :, 11:].abs().max() == 0
    back = torch.zeros(4, 3, 100, device=DEV)
    _lib.call("b2d_cl_to_planar", y.data_ptr(), None, back.data_ptr(), 4, 3, 100, 64, 8, 0, _s())
    assert torch.equal(back, ref.float())
    # the same through IEEE fp16 storage
    yh = torch.zeros(4, 100, 64, dtype=torch.float16, device=DEV)
    _lib.call("b2d_planar_to_cl", x.data_ptr(), yh.data_ptr(), None, 4, 3, 100, 64, 8, scale.data_ptr(), 1, _s())
    refh = (x / scale.view(1, 3, 1)).to(torch.float16)
    assert torch.equal(yh[:, :, 8:11], refh.permute(0, 2, 1)) and yh[:, :, :8].abs().max() == 0 and yh[:, :, 11:].abs().max() == 0
    _lib.call("b2d_cl_to_planar", yh.data_ptr(), None, back.data_ptr(), 4, 3, 100, 64, 8, 1, _s())
    assert torch.equal(back, refh.float())


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("N,T,C", [(3, 256, 256), (5, 64, 512), (4, 16, 1024), (2, 4, 256), (2, 1, 1024)])
def test_attention_core(N, T, C, f16):
    """T >= 16: the tcgen05 kernel; T < 16: the CUDA-core kernel.  Both in bf16 and in IEEE fp16 storage."""
    no_tf32()
    heads = 2
    dt = torch.float16 if f16 else torch.bfloat16
    g = torch.Generator().manual_seed(T + C)
    qkv = fmt_round(torch.randn(N, T, 3 * C, generator=g).to(DEV), f16)
    out = torch.zeros(N, T, C, dtype=dt, device=DEV)
    _lib.call("b2d_attention", qkv.to(dt).contiguous().data_ptr(), None, out.data_ptr(), None, N, T, C, heads, 1 if f16 else 0, _s())
    q, k, v = qkv.chunk(3, dim=-1)
    d = C // heads
    sp = lambda t: t.reshape(N, T, heads, d).transpose(1, 2)
    att = torch.softmax((sp(q) / math.sqrt(d)) @ sp(k).transpose(-1, -2), dim=-1)
    ref = (att @ sp(v)).transpose(1, 2).reshape(N, T, C)
    assert rel_err(out, ref) < (2e-3 if f16 else 1e-2)


def test_edt_exact_vs_scipy_and_bilinear():
    from scipy import ndimage
    g = torch.Generator().manual_seed(4)
    imgs = (torch.rand(5, 64, 96, generator=g) > 0.35).float()
    imgs[1, :, :40] = 1.0  # a large all-foreground region
    imgs[2] = 0.0          # all background
    out = torch.zeros(2, 5, 64, 96, device=DEV)
    _lib.call("b2d_edt2d", imgs.to(DEV).data_ptr(), out.data_ptr(), 5, 64, 96, _s(), launches=2)
    ref = torch.from_numpy(np.stack([ndimage.distance_transform_edt(im.numpy()) for im in imgs])).float()
    assert torch.equal(out[0].cpu(), ref)  # predictor.py:1096-1116
    small = torch.zeros(5, 16, 24, device=DEV)
    _lib.call("b2d_bilinear_resize", out[0].data_ptr(), small.data_ptr(), 5, 64, 96, 16, 24, _s())
    refb = F.interpolate(ref[:, None], size=(16, 24), mode="bilinear", align_corners=False)[:, 0]
    assert (small.cpu() - refb).abs().max().item() <= 1e-5 * refb.abs().max().item()
    up = torch.zeros(5, 100, 50, device=DEV)
    _lib.call("b2d_bilinear_resize", out[0].data_ptr(), up.data_ptr(), 5, 64, 96, 100, 50, _s())
    refu = F.interpolate(ref[:, None], size=(100, 50), mode="bilinear", align_corners=False)[:, 0]
    assert (up.cpu() - refu).abs().max().item() <= 1e-5 * refu.abs().max().item()
