"""Per-sample fused GroupNorm launches (b2d_gn_gn_apply, b2d_maxpool2x2_gn) against the two-launch forms they replace
(b2d_gn_apply with stats_out + b2d_gn_apply; b2d_maxpool2x2_stats + b2d_gn_apply) and against plain fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _acts(N, H, W, C, seed, f16=False):
    from diffusion_model_project_b200.engine import new_act
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(N, 1, H, W, C, generator=g) * 1.7 + 0.3
    a = new_act(N, 1, H, W, C, "cuda", f16=f16)
    if f16:
        a.hi.view(torch.float16).copy_(v.to(torch.float16))
        v = v.to(torch.float16).float()
    else:
        a.hi.copy_(v.to(torch.bfloat16))
        v = v.to(torch.bfloat16).float()
    return a, v


def _val(a):
    return a.hi.view(torch.float16).float() if a.f16 else a.hi.float()


@pytest.mark.parametrize("of16", [False, True])
@pytest.mark.parametrize("H,C", [(16, 256), (8, 512), (4, 1024), (2, 2048), (32, 64)])
def test_gn_gn_apply(H, C, of16):
    """of16: the normalised outputs are stored as IEEE fp16 (the default 16-bit mode) instead of bf16."""
    from diffusion_model_project_b200 import engine
    from diffusion_model_project_b200.engine import new_act as _new_act
    new_act = lambda *a: _new_act(*a, f16=of16)
    tol = 3e-3 if of16 else 2e-2
    N = 5
    x, xv = _acts(N, H, H, C, seed=H, f16=True)
    g = torch.Generator().manual_seed(1)
    g1, b1, g2, b2 = (torch.randn(C, generator=g).cuda() for _ in range(4))
    st = torch.stack([xv.double().sum(dim=(1, 2, 3, 4)), (xv.double() ** 2).sum(dim=(1, 2, 3, 4))], dim=1).cuda().contiguous()
    s = torch.cuda.current_stream().cuda_stream
    # reference: fp32 torch on channels-first
    xc = xv[:, 0].permute(0, 3, 1, 2).cuda()
    r1 = F.silu(F.group_norm(xc, 1, g1, b1, 1e-5))
    r2 = F.group_norm(r1, 1, g2, b2, 1e-5)
    y1, y2 = new_act(N, 1, H, H, C, "cuda"), new_act(N, 1, H, H, C, "cuda")
    engine.gn_gn_apply(x, y1, y2, st, g1, b1, True, g2, b2, False, s)
    torch.cuda.synchronize()
    for got, ref in ((y1, r1), (y2, r2)):
        got = _val(got)[:, 0].permute(0, 3, 1, 2)
        assert (got - ref).abs().max() <= tol * ref.abs().max()
    # the two-launch form it replaces
    z1, z2 = new_act(N, 1, H, H, C, "cuda"), new_act(N, 1, H, H, C, "cuda")
    st2 = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    engine.gn_apply(x, z1, st, C, g1, b1, True, s, stats_out=st2)
    engine.gn_apply(z1, z2, st2, C, g2, b2, False, s)
    torch.cuda.synchronize()
    assert (_val(z1) - _val(y1)).abs().max() <= tol / 2 * _val(z1).abs().max()  # same math, 0.5 folded differently
    assert (_val(z2) - _val(y2)).abs().max() <= tol * _val(z2).abs().max()


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("H,C", [(64, 64), (32, 128), (16, 256), (8, 512), (4, 1024)])
def test_maxpool_gn(H, C, f16):
    from diffusion_model_project_b200 import engine
    from diffusion_model_project_b200.engine import new_act as _new_act
    new_act = lambda *a: _new_act(*a, f16=f16)
    tol = 3e-3 if f16 else 2e-2
    N = 3
    x, xv = _acts(N, H, H, C, seed=H + 1, f16=f16)
    g = torch.Generator().manual_seed(2)
    ga, be = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    s = torch.cuda.current_stream().cuda_stream
    ref = F.silu(F.group_norm(F.max_pool2d(xv[:, 0].permute(0, 3, 1, 2).cuda(), 2), 1, ga, be, 1e-5))
    y = new_act(N, 1, H // 2, H // 2, C, "cuda")
    engine.maxpool_gn(x, y, ga, be, True, s)
    torch.cuda.synchronize()
    got = _val(y)[:, 0].permute(0, 3, 1, 2)
    assert (got - ref).abs().max() <= tol * ref.abs().max()
    p = new_act(N, 1, H // 2, H // 2, C, "cuda")
    st = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    engine.maxpool_stats(x, p, st, s)
    engine.gn_apply(p, p, st, C, ga, be, True, s)
    torch.cuda.synchronize()
    assert (_val(p) - _val(y)).abs().max() <= tol / 2 * _val(p).abs().max()


def test_fused_ops_reject_large_samples():
    from diffusion_model_project_b200 import engine
    from diffusion_model_project_b200._lib import B2DError
    from diffusion_model_project_b200.engine import new_act
    x = new_act(1, 1, 64, 64, 64, "cuda")
    y = new_act(1, 1, 64, 64, 64, "cuda")
    st = torch.zeros(1, 2, dtype=torch.float64, device="cuda")
    with pytest.raises((B2DError, RuntimeError, ValueError)):
        engine.gn_gn_apply(x, y, y, st, None, None, True, None, None, False, torch.cuda.current_stream().cuda_stream)
