"""tcgen05 implicit-GEMM conv engine vs plain PyTorch fp32 (TF32 off) on bf16-rounded operands.
Tolerance: the engine multiplies exact bf16 products and accumulates in fp32, so against an fp32
reference on the same rounded operands max|err|/max|ref| <= 1e-2 is dominated by the bf16 rounding of
the stored output (2^-9); fp32 outputs and the fp32x (hi/lo split) mode are held to 2e-4."""
import pytest
import torch
import torch.nn.functional as F

from diffusion_model_project_b200 import _lib, engine
from diffusion_model_project_b200.engine import ConvPlan, new_act
from util import bf16_round, f16_round, from_act, no_tf32, rel_err, stats_ref, to_act

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"
TOL_BF16 = 1e-2
TOL_F32 = 2e-4


def _rnd(g, *shape, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("bn,cout,K", [(16, 8, 64), (16, 3, 128), (64, 64, 64), (64, 192, 256), (128, 256, 512), (256, 512, 192)])
def test_gemm_block_n_variants(bn, cout, K):
    no_tf32()
    g = torch.Generator().manual_seed(bn + K)
    x = bf16_round(_rnd(g, 3, K, 1, 16, 16))
    w = bf16_round(_rnd(g, cout, K, scale=K ** -0.5))
    b = _rnd(g, cout)
    pw = engine.pack_linear(w, b, DEV)
    out = torch.zeros(3, 16, 16, max(cout, 4), device=DEV)
    plan = ConvPlan([to_act(x)], pw, out, cout=cout, out_mode=2, out_cstride=out.shape[-1], block_n=bn)
    assert plan.info()["block_n"] == bn
    plan.run(_stream())
    ref = F.conv2d(x[:, :, 0], w[:, :, None, None], b)
    assert rel_err(out[..., :cout].permute(0, 3, 1, 2), ref) < TOL_F32


@pytest.mark.parametrize("shape", [(3, 64, 128, 64, 64), (11, 128, 128, 4, 4), (5, 64, 64, 2, 2), (7, 64, 64, 1, 1),
                                   (3, 256, 64, 8, 8), (2, 17, 64, 32, 32), (22, 512, 128, 4, 4), (1, 64, 64, 16, 48)])
def test_conv2d_3x3_zero_padding_and_gn_sums(shape):
    no_tf32()
    N, ci, co, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = bf16_round(_rnd(g, N, ci, 1, H, W))
    w = bf16_round(_rnd(g, co, ci, 3, 3, scale=(9 * ci) ** -0.5))
    pw = engine.pack_conv2d(w, [ci], None, DEV)
    out = new_act(N, 1, H, W, co, DEV)
    st = torch.zeros(N, 1, 2, dtype=torch.float64, device=DEV)
    ConvPlan([to_act(x)], pw, out, cout=co, stats=st, stats_cpg=co).run(_stream())
    ref = F.conv2d(x[:, :, 0], w, None, padding=1)
    assert rel_err(from_act(out, co)[:, :, 0], ref) < TOL_BF16
    sref = stats_ref(ref[:, :, None], 1)
    assert ((st - sref).abs().max() / sref.abs().max()).item() < 1e-4


@pytest.mark.parametrize("kind,shape", [("3x3", (3, 64, 128, 32, 32)), ("3x3", (5, 256, 128, 8, 8)), ("3x3", (7, 1024, 512, 2, 2)),
                                        ("1x1", (3, 256, 768, 16, 16)), ("convT", (3, 128, 64, 8, 8)), ("3d", (2, 128, 128, 16, 16))])
def test_fp16_operands(kind, shape):
    """b2d_conv_desc.op_f16: activations and weights stored as IEEE fp16 (the default 16-bit mode) through the same
    kernels -- halo and generic staging, split-K, transposed-conv phases, 3x3x3 -- with fp16 and bf16 outputs, residual
    and GroupNorm sums.  With fp16-rounded operands the only error left is the 16-bit rounding of the OUTPUT."""
    no_tf32()
    N, ci, co, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) + len(kind))
    D = 3 if kind == "3d" else 1
    x = f16_round(_rnd(g, N, ci, D, H, W))
    b = _rnd(g, co) if kind != "3x3" else None
    if kind == "3x3":
        w = f16_round(_rnd(g, co, ci, 3, 3, scale=(9 * ci) ** -0.5))
        pw = engine.pack_conv2d(w, [ci], None, DEV, f16=True)
        ref = F.conv2d(x[:, :, 0], w, None, padding=1)[:, :, None]
    elif kind == "1x1":
        w = f16_round(_rnd(g, co, ci, scale=ci ** -0.5))
        pw = engine.pack_linear(w, b, DEV, f16=True)
        ref = F.conv2d(x[:, :, 0], w[:, :, None, None], b)[:, :, None]
    elif kind == "convT":
        w = f16_round(_rnd(g, ci, co, 2, 2, scale=ci ** -0.5))
        pw = engine.pack_convT2x2(w, b, DEV, f16=True)
        ref = F.conv_transpose2d(x[:, :, 0], w, b, stride=2)[:, :, None]
    else:
        w = f16_round(_rnd(g, co, ci, 3, 3, 3, scale=(27 * ci) ** -0.5))
        pw = engine.pack_conv3d(w, b, DEV, f16=True)
        ref = F.conv3d(x, w, b, padding=1)
    assert pw.f16
    up = 2 if kind == "convT" else 1
    res = f16_round(_rnd(g, *ref.shape))
    for out_f16 in (True, False):
        out = new_act(N, D, H * up, W * up, co, DEV, f16=out_f16)
        st = torch.zeros(N, 1, 2, dtype=torch.float64, device=DEV)
        plan = ConvPlan([to_act(x, f16=True)], pw, out, cout=co, nphase=4 if kind == "convT" else 1, stats=st, stats_cpg=co,
                        residual=to_act(res, f16=True))
        plan.run(_stream())
        want = ref + res
        assert rel_err(from_act(out, co), want) < (1.5e-3 if out_f16 else TOL_BF16)
        sref = stats_ref(want, 1)
        assert ((st - sref).abs().max() / sref.abs().max()).item() < 1e-4
    with pytest.raises(AssertionError):   # operands and weights must share one 16-bit format
        ConvPlan([to_act(x)], pw, new_act(N, D, H * up, W * up, co, DEV), cout=co, nphase=4 if kind == "convT" else 1)


def test_concat_as_two_k_segments():
    no_tf32()
    g = torch.Generator().manual_seed(3)
    N, c, H, W = 3, 128, 16, 16
    skip, up = bf16_round(_rnd(g, N, c, 1, H, W)), bf16_round(_rnd(g, N, c, 1, H, W))
    w = bf16_round(_rnd(g, c, 2 * c, 3, 3, scale=(18 * c) ** -0.5))
    pw = engine.pack_conv2d(w, [c, c], None, DEV)
    out = new_act(N, 1, H, W, c, DEV)
    ConvPlan([to_act(skip), to_act(up)], pw, out, cout=c).run(_stream())
    ref = F.conv2d(torch.cat((skip, up), 1)[:, :, 0], w, None, padding=1)  # unet/models.py:177
    assert rel_err(from_act(out, c)[:, :, 0], ref) < TOL_BF16


@pytest.mark.parametrize("shape", [(3, 128, 64, 8, 8), (11, 2048, 1024, 2, 2), (2, 128, 64, 32, 32), (5, 256, 128, 1, 1)])
def test_conv_transpose_2x2(shape):
    no_tf32()
    N, ci, co, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = bf16_round(_rnd(g, N, ci, 1, H, W))
    w = bf16_round(_rnd(g, ci, co, 2, 2, scale=ci ** -0.5))
    b = _rnd(g, co)
    pw = engine.pack_convT2x2(w, b, DEV)
    out = new_act(N, 1, 2 * H, 2 * W, co, DEV)
    st = torch.zeros(N, 1, 2, dtype=torch.float64, device=DEV)
    ConvPlan([to_act(x)], pw, out, cout=co, nphase=4, stats=st, stats_cpg=co).run(_stream())
    ref = F.conv_transpose2d(x[:, :, 0], w, b, stride=2)
    assert rel_err(from_act(out, co)[:, :, 0], ref) < TOL_BF16
    sref = stats_ref(ref[:, :, None], 1)
    assert ((st - sref).abs().max() / sref.abs().max()).item() < 1e-4


@pytest.mark.parametrize("down", [False, True])
@pytest.mark.parametrize("shape", [(2, 128, 128, 3, 16, 16), (1, 64, 256, 11, 8, 8), (2, 128, 128, 3, 4, 4), (1, 3, 128, 2, 32, 32),
                                   (1, 512, 512, 2, 8, 8)])
def test_conv3d_and_strided_down(shape, down):
    no_tf32()
    N, ci, co, D, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) + down)
    x = bf16_round(_rnd(g, N, ci, D, H, W))
    w = bf16_round(_rnd(g, co, ci, 3, 3, 3, scale=(27 * ci) ** -0.5))
    b = _rnd(g, co)
    pw = engine.pack_conv3d(w, b, DEV, down=down)
    s_ = 2 if down else 1
    out = new_act(N, D, H // s_, W // s_, co, DEV)
    st = torch.zeros(N, 32, 2, dtype=torch.float64, device=DEV)
    ConvPlan([to_act(x)], pw, out, cout=co, stride=s_, stats=st, stats_cpg=co // 32).run(_stream())
    ref = F.conv3d(F.pad(x, (0, 1, 0, 1, 1, 1)), w, b, stride=(1, 2, 2)) if down else F.conv3d(x, w, b, padding=1)
    assert rel_err(from_act(out, co), ref) < TOL_BF16
    sref = stats_ref(ref, 32)
    assert ((st - sref).abs().max() / sref.abs().max()).item() < 1e-4


def test_residual_add_and_1x1x1():
    no_tf32()
    g = torch.Generator().manual_seed(11)
    N, ci, co, D, H, W = 2, 128, 256, 2, 8, 8
    x = bf16_round(_rnd(g, N, ci, D, H, W))
    r = bf16_round(_rnd(g, N, co, D, H, W))
    w = bf16_round(_rnd(g, co, ci, 1, 1, 1, scale=ci ** -0.5))
    b = _rnd(g, co)
    pw = engine.pack_conv3d(w, b, DEV)
    out = new_act(N, D, H, W, co, DEV)
    ConvPlan([to_act(x)], pw, out, cout=co, residual=to_act(r)).run(_stream())
    assert rel_err(from_act(out, co), F.conv3d(x, w, b) + r) < TOL_BF16


def test_planar_output_with_scale_and_mask():
    no_tf32()
    g = torch.Generator().manual_seed(12)
    N, ci, co, D, H, W = 2, 128, 3, 2, 16, 16
    x = bf16_round(_rnd(g, N, ci, D, H, W))
    w = bf16_round(_rnd(g, co, ci, 3, 3, 3, scale=(27 * ci) ** -0.5))
    b = _rnd(g, co)
    scale = torch.tensor([0.01, 0.005, 0.002], device=DEV)
    mask = (torch.rand(N, D, H, W, generator=g) > 0.4).float().to(DEV)
    out = torch.zeros(N, D, co, H, W, device=DEV)
    ConvPlan([to_act(x)], engine.pack_conv3d(w, b, DEV), out, cout=co, out_mode=1, out_cstride=co, out_scale=scale,
             out_mask=mask).run(_stream())
    ref = F.conv3d(x, w, b, padding=1) * scale.view(1, 3, 1, 1, 1) * mask[:, None]
    assert rel_err(out.permute(0, 2, 1, 3, 4), ref) < TOL_F32


def test_channel_offset_output():
    """E2D conv_out writes mu straight into channels [8,16) of the UNet input buffer."""
    no_tf32()
    g = torch.Generator().manual_seed(13)
    N, ci, D, H, W = 1, 512, 2, 8, 8
    x = bf16_round(_rnd(g, N, ci, D, H, W))
    w = bf16_round(_rnd(g, 16, ci, 3, 3, 3, scale=(27 * ci) ** -0.5))
    b = _rnd(g, 16)
    out = new_act(N, D, H, W, 64, DEV, zero=True)
    out.hi.fill_(7.0)
    ConvPlan([to_act(x)], engine.pack_conv3d(w, b, DEV), out, cout=8, out_coff=8).run(_stream())
    ref = F.conv3d(x, w, b, padding=1)[:, :8]
    got = out.hi.float()
    assert rel_err(got[..., 8:16].permute(0, 4, 1, 2, 3), ref) < TOL_BF16
    assert (got[..., :8] == 7).all() and (got[..., 16:] == 7).all()


def test_fp32x_split_mode():
    no_tf32()
    g = torch.Generator().manual_seed(14)
    N, ci, co, H, W = 2, 128, 128, 16, 16
    x = _rnd(g, N, ci, 1, H, W)
    w = _rnd(g, co, ci, 3, 3, scale=(9 * ci) ** -0.5)
    pw = engine.pack_conv2d(w, [ci], None, DEV, split=True)
    out = new_act(N, 1, H, W, co, DEV, split=True)
    ConvPlan([to_act(x, split=True)], pw, out, cout=co).run(_stream())
    ref = F.conv2d(x[:, :, 0].double(), w.double(), None, padding=1).float()
    assert rel_err(from_act(out, co)[:, :, 0], ref) < 1e-4


def test_plan_rejects_bad_shapes():
    x = new_act(1, 1, 8, 8, 64, DEV, zero=True)
    w = torch.zeros(60, 64)
    with pytest.raises(ValueError):
        ConvPlan([x], engine.pack_linear(w, None, DEV), new_act(1, 1, 8, 8, 64, DEV), cout=60)  # cout not multiple of 64


@pytest.mark.parametrize("ci,co,f16", [(128, 128, True), (256, 128, False), (512, 16, True)])
def test_fused_input_groupnorm_silu(ci, co, f16):
    """conv(silu(GroupNorm32(x))) with the normalisation applied to the staged tiles inside the conv kernel
    (vae/blocks.py:173-183), raw input held as fp16 or bf16, vs PyTorch on the same rounded operands."""
    no_tf32()
    g = torch.Generator().manual_seed(21 + ci)
    N, D, H, W = 2, 3, 32, 16
    x = _rnd(g, N, ci, D, H, W) * 1.5 + 0.3
    xr = x.to(torch.float16).float() if f16 else bf16_round(x)
    gamma = (1 + 0.1 * torch.randn(ci, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(ci, generator=g)).to(DEV)
    cout_real = co if co >= 64 else 3
    w = bf16_round(_rnd(g, cout_real, ci, 3, 3, 3, scale=(27 * ci) ** -0.5))
    b = _rnd(g, cout_real)
    if f16:  # the same 16-bit storage, holding IEEE half values
        cl = xr.permute(0, 2, 3, 4, 1).to(torch.float16).contiguous()
        xa = engine.Act(cl.view(torch.bfloat16), None, True)
    else:
        xa = to_act(xr)
    st = stats_ref(xr, 32).to(DEV).contiguous()
    if cout_real >= 64:
        out = new_act(N, D, H, W, co, DEV)
        plan = ConvPlan([xa], engine.pack_conv3d(w, b, DEV), out, cout=co, in_norm=(st, ci // 32, gamma, beta, True))
        plan.run(_stream())
        got = from_act(out, co)
    else:
        out = torch.zeros(N, D, cout_real, H, W, device=DEV)
        plan = ConvPlan([xa], engine.pack_conv3d(w, b, DEV), out, cout=cout_real, out_mode=1, out_cstride=cout_real,
                        in_norm=(st, ci // 32, gamma, beta, True))
        plan.run(_stream())
        got = out.permute(0, 2, 1, 3, 4)
    assert plan.info2()["halo"] == 1
    h = bf16_round(F.silu(F.group_norm(xr, 32, gamma, beta, eps=1e-5)))
    ref = F.conv3d(h, w, b, padding=1)
    assert rel_err(got, ref) < (TOL_BF16 if cout_real >= 64 else 2e-3)


def _conv2d_case(g, N, ci, co, H, W):
    x = bf16_round(_rnd(g, N, ci, 1, H, W))
    w = bf16_round(_rnd(g, co, ci, 3, 3, scale=(9 * ci) ** -0.5))
    return x, w


@pytest.mark.parametrize("N,ci,co,H,W", [(3, 64, 64, 32, 16), (2, 128, 256, 16, 16), (5, 256, 128, 8, 8), (7, 512, 512, 4, 4)])
def test_tiling_variants_agree(N, ci, co, H, W):
    """The plan's own choices (halo staging where W,H % 16 == 0, split-K where the tile count is small, contiguous unit
    walk) against the plain comparison arm selected through b2d_conv_desc.tune_flags (generic 128-row tiles, no K split,
    strided walk) on the same operands: both accumulate bf16 products in fp32, only the summation order differs."""
    no_tf32()
    g = torch.Generator().manual_seed(31 + ci)
    x, w = _conv2d_case(g, N, ci, co, H, W)
    pw = engine.pack_conv2d(w, [ci], None, DEV)
    outs, stats, infos = [], [], []
    plain = _lib.TUNE_NO_HALO | _lib.TUNE_NO_SPLITK | _lib.TUNE_STRIDED
    for flags in (plain, 0):
        out = new_act(N, 1, H, W, co, DEV)
        st = torch.zeros(N, 2, dtype=torch.float64, device=DEV)
        plan = ConvPlan([to_act(x)], pw, out, cout=co, stats=st, stats_cpg=co, tune_flags=flags)
        plan.run(_stream())
        outs.append(from_act(out, co)); stats.append(st.clone()); infos.append(plan.info2())
    assert infos[0]["halo"] == 0 and infos[0]["ksplit"] == 1
    assert infos[1]["halo"] == (1 if H % 16 == 0 and W % 16 == 0 else 0)
    ref = F.conv2d(x[:, :, 0], w, None, padding=1)[:, :, None]
    assert rel_err(outs[1], ref) < TOL_BF16
    assert rel_err(outs[1], outs[0]) < 8e-3  # one bf16 ulp of the output at most
    assert ((stats[1] - stats[0]).abs().max() / stats[0].abs().max()).item() < 1e-5


def test_forced_k_split_matches_cost_model_choice():
    """b2d_conv_desc.tune_ksplit (tools/tune_conv.py) forces a split count; the result does not depend on it beyond fp32
    summation order."""
    no_tf32()
    g = torch.Generator().manual_seed(43)
    N, ci, co = 5, 512, 512
    x, w = _conv2d_case(g, N, ci, co, 4, 4)
    pw = engine.pack_conv2d(w, [ci], None, DEV)
    outs = []
    for ks in (1, 3, 8):
        out = new_act(N, 1, 4, 4, co, DEV)
        plan = ConvPlan([to_act(x)], pw, out, cout=co, tune_ksplit=ks)
        assert plan.info2()["ksplit"] == ks
        plan.run(_stream())
        outs.append(from_act(out, co))
    assert rel_err(outs[1], outs[0]) < 8e-3 and rel_err(outs[2], outs[0]) < 8e-3


@pytest.mark.parametrize("N,ci,co,H,W,f16", [(88, 256, 256, 16, 16, True), (88, 128, 256, 16, 16, False), (88, 512, 512, 8, 8, True),
                                              (88, 1024, 1024, 4, 4, True), (88, 2048, 2048, 2, 2, False), (40, 256, 512, 8, 8, True),
                                              (7, 512, 1024, 4, 4, True), (88, 128, 128, 32, 32, True)])
def test_stream_k_matches_unit_walk(N, ci, co, H, W, f16):
    """Stream-K (csrc/conv_work.cuh: the flat (tile, K step) space cut into one equal range per CTA, shared tiles reduced
    through the workspace in piece order) forced on, against the plain unit walk of the same plan: UNet layer shapes at
    88 slice-images, halo and generic staging, GroupNorm sums.  Repeated launches are bit-identical (deterministic
    reduction order) and the arrival counters return to zero."""
    no_tf32()
    g = torch.Generator().manual_seed(ci + co + H)
    rnd = f16_round if f16 else bf16_round
    x = rnd(_rnd(g, N, ci, 1, H, W))
    w = rnd(_rnd(g, co, ci, 3, 3, scale=(9 * ci) ** -0.5))
    pw = engine.pack_conv2d(w, [ci], None, DEV, f16=f16)
    xa = to_act(x, f16=f16)
    res = {}
    for name, flags in (("unit", _lib.TUNE_NO_STREAMK | _lib.TUNE_NO_SPLITK), ("streamk", _lib.TUNE_STREAMK)):
        ws = engine.new_workspace(DEV, 128 << 20)
        out = new_act(N, 1, H, W, co, DEV, f16=True)
        st = torch.zeros(N, 2, dtype=torch.float64, device=DEV)
        plan = ConvPlan([xa], pw, out, cout=co, stats=st, stats_cpg=co, workspace=ws, tune_flags=flags)
        info = plan.info2()
        plan.run(_stream())
        first, st_first = out.hi.clone(), st.clone()
        if name == "streamk":
            assert info["ksplit"] == -1 and info["ctas"] == torch.cuda.get_device_properties(0).multi_processor_count, info
            for _ in range(2):
                out.hi.zero_(); st.zero_()
                plan.run(_stream())
                assert torch.equal(out.hi, first)
            assert int(ws[:16384].view(torch.int32).abs().sum()) == 0
        res[name] = (from_act(out, co), st_first)
    ref = F.conv2d(x[:, :, 0], w, None, padding=1)[:, :, None]
    assert rel_err(res["streamk"][0], ref) < 1.5e-3                      # fp16 output rounding
    assert rel_err(res["streamk"][0], res["unit"][0]) < 1.5e-3           # fp32 summation order only
    assert ((res["streamk"][1] - res["unit"][1]).abs().max() / res["unit"][1].abs().max()).item() < 5e-5  # sums of fp16-rounded values


@pytest.mark.parametrize("kind,N,ci,co,H,W", [("3x3", 88, 512, 512, 8, 8), ("3x3", 88, 1024, 1024, 4, 4), ("3x3", 88, 2048, 2048, 2, 2),
                                              ("3x3", 24, 512, 256, 4, 4), ("3x3", 5, 256, 512, 8, 8), ("1x1", 88, 256, 768, 16, 16),
                                              ("1x1", 11, 1024, 1024, 4, 4)])
def test_cta_pairs_match_single_cta_tiles(kind, N, ci, co, H, W):
    """tcgen05 cta_group::2 (B2D_TUNE_PAIR: clusters of two CTAs on 256-row M super-tiles, each staging half of the weight
    rows) against the one-CTA-per-tile plan of the same layer: deep UNet level shapes at 88 slice-images with and without
    split-K, an odd number of M tiles (the pair's second tile is all padding), bias + residual, GroupNorm sums.  Repeated
    launches are bit-identical."""
    no_tf32()
    g = torch.Generator().manual_seed(ci + co + H + N)
    x = f16_round(_rnd(g, N, ci, 1, H, W))
    if kind == "3x3":
        w = f16_round(_rnd(g, co, ci, 3, 3, scale=(9 * ci) ** -0.5))
        pw = engine.pack_conv2d(w, [ci], None, DEV, f16=True)
        ref = F.conv2d(x[:, :, 0], w, None, padding=1)[:, :, None]
        resid = None
    else:
        w = f16_round(_rnd(g, co, ci, scale=ci ** -0.5))
        b = _rnd(g, co)
        pw = engine.pack_linear(w, b, DEV, f16=True)
        resid = f16_round(_rnd(g, N, co, 1, H, W))
        ref = F.conv2d(x[:, :, 0], w[:, :, None, None], b)[:, :, None] + resid
    xa = to_act(x, f16=True)
    res = {}
    for name, flags in (("single", _lib.TUNE_NO_PAIR), ("pair", _lib.TUNE_PAIR)):
        out = new_act(N, 1, H, W, co, DEV, f16=True)
        st = torch.zeros(N, 2, dtype=torch.float64, device=DEV)
        ws = engine.new_workspace(DEV)
        plan = ConvPlan([xa], pw, out, cout=co, stats=st, stats_cpg=co, workspace=ws, tune_flags=flags,
                        residual=None if resid is None else to_act(resid, f16=True))
        info = plan.info2()
        assert info["engine"] == (3 if name == "pair" else 2), info
        plan.run(_stream())
        first, st_first = out.hi.clone(), st.clone()
        for _ in range(2):
            out.hi.zero_(); st.zero_()
            plan.run(_stream())
            assert torch.equal(out.hi, first)
        assert int(ws[:16384].view(torch.int32).abs().sum()) == 0
        res[name] = (from_act(out, co), st_first)
    assert rel_err(res["pair"][0], ref) < 1.5e-3
    assert rel_err(res["pair"][0], res["single"][0]) < 1.5e-3
    assert ((res["pair"][1] - res["single"][1]).abs().max() / res["single"][1].abs().max()).item() < 5e-5


def test_split_k_is_deterministic_and_reuses_workspace():
    """Deep UNet level shape (M = 7*4 rows, K = 9*2048): the plan splits K, partials go through the shared workspace,
    the last arriver reduces in a fixed order -> repeated launches are bit-identical and the counters self-reset."""
    no_tf32()
    g = torch.Generator().manual_seed(41)
    N, ci, co = 7, 2048, 1024
    x, w = _conv2d_case(g, N, ci, co, 2, 2)
    pw = engine.pack_conv2d(w, [ci], None, DEV)
    out = new_act(N, 1, 2, 2, co, DEV)
    st = torch.zeros(N, 2, dtype=torch.float64, device=DEV)
    ws = engine.new_workspace(DEV)
    plan = ConvPlan([to_act(x)], pw, out, cout=co, stats=st, stats_cpg=co, workspace=ws)
    info = plan.info2()
    assert info["ksplit"] > 1 and info["ws_kib"] > 0, info
    plan.run(_stream())
    first = out.hi.clone()
    for _ in range(3):
        out.hi.zero_()
        plan.run(_stream())
        assert torch.equal(out.hi, first)
    ref = F.conv2d(x[:, :, 0], w, None, padding=1)[:, :, None]
    assert rel_err(from_act(out, co), ref) < TOL_BF16
    assert int(ws[:16384].view(torch.int32).abs().sum()) == 0  # arrival counters back to zero


@pytest.mark.parametrize("ci,co", [(256, 128), (128, 256)])
def test_upsample_folded_conv_matches_upsample_then_conv(ci, co):
    """decoder.py:46-47 / 58-59: Conv3d(nn.Upsample(scale=(1,2,2))(x)) as four phase convs on the low-res map."""
    no_tf32()
    g = torch.Generator().manual_seed(51 + ci)
    N, D, H, W = 2, 3, 16, 32
    x = bf16_round(_rnd(g, N, ci, D, H, W))
    w = _rnd(g, co, ci, 3, 3, 3, scale=(27 * ci) ** -0.5).cpu()
    b = _rnd(g, co).cpu()
    out = new_act(N, D, 2 * H, 2 * W, co, DEV)
    st = torch.zeros(N, 32, 2, dtype=torch.float64, device=DEV)
    xa = to_act(x)
    for ph in range(4):
        pw = engine.pack_conv3d_upsampled(w, b, DEV, ph >> 1, ph & 1)
        plan = ConvPlan([xa], pw, out, cout=co, stats=st, stats_cpg=co // 32, out_geom=(2 * H, 2 * W, 2, 2, ph >> 1, ph & 1))
        assert plan.info2()["halo"] == 1 and plan.info2()["kgroups"] == 3 * ci // 64
        plan.run(_stream())
    ref = F.conv3d(F.interpolate(x, scale_factor=(1, 2, 2), mode="nearest"), w.to(DEV), b.to(DEV), padding=1)
    assert rel_err(from_act(out, co), ref) < 1.5e-2  # summed-then-rounded weights vs rounded-then-summed
    sref = stats_ref(ref, 32)
    assert ((st - sref).abs().max() / sref.abs().max()).item() < 2e-2
