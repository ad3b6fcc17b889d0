"""Training-step slice on the GPU: the criterion, Adam, and the backward of one DoubleBlock
(GroupNorm/SiLU backward, conv data gradient through the forward engine, tcgen05 weight gradient) against torch autograd
through the oracle's functional UNet blocks and against the golden vectors of one training step of the unmodified
reference (tests/golden/train_step.npz).  Bounds: loss <= 1e-5, gradients rel-L2 <= 2e-3 (fp32-class mode)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffusion_model_project_b200 import _lib, engine, synth, train
from diffusion_model_project_b200.engine import new_act
from oracle import train as otrain
from oracle import unet as ounet
from util import f16_round, from_act, no_tf32, rel_err, rel_l2, to_act

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _autograd_on():
    """Other test modules switch autograd off process-wide at import; the references here are torch.autograd results."""
    with torch.enable_grad():
        yield


def _s():
    return torch.cuda.current_stream().cuda_stream


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "train_step.npz"))


def test_nmse_loss_forward_and_gradient(golden_dir):
    g = torch.Generator().manual_seed(3)
    for shape, weight in (((4, 8, 32, 32), None), ((3, 5, 7, 9), torch.tensor([1.0, 2.0, 0.5, 1.5, 3.0])), ((2, 3, 4, 8, 8), None)):
        o = torch.randn(*shape, generator=g)
        t = torch.randn(*shape, generator=g) * 0.7
        oo = o.clone().requires_grad_(True)
        ref = otrain.normalized_mse_loss_per_component(oo, t, weight_per_channel=weight)
        ref.backward()
        per = otrain.normalized_mse_loss_per_component(o, t, reduce=False, weight_per_channel=weight)
        loss, per_gpu, grad = train.nmse_loss(o.to(DEV), t.to(DEV), None if weight is None else weight.to(DEV))
        assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
        assert rel_err(per_gpu.cpu(), per) <= 1e-5
        assert rel_err(grad.cpu(), oo.grad) <= 1e-5
    with torch.no_grad():
        with pytest.raises(ValueError):
            train.nmse_loss(torch.zeros(2, 3, 4, device=DEV), torch.zeros(2, 3, 4, device=DEV))   # metrics.py:369
    # the reference's own training step: its eps-prediction and target noise give its loss
    gd = _golden(golden_dir)
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_train_golden", os.path.join(golden_dir, "make_train_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    import sys
    saved = list(sys.path)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved
    _, _, _, noise, _ = mod.train_inputs()
    loss, _, _ = train.nmse_loss(torch.from_numpy(gd["pred"]).to(DEV), noise.to(DEV), want_grad=False)
    assert abs(loss.item() - float(gd["loss"])) <= 1e-5 * abs(float(gd["loss"]))


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_flat_adam_matches_torch_adam(wd, golden_dir):
    g = torch.Generator().manual_seed(5)
    params = {"a.weight": torch.randn(64, 17, 3, 3, generator=g) * 0.1, "a.bias": torch.randn(7, generator=g), "b": torch.randn(1001, generator=g)}
    opt = train.FlatAdam(params, lr=1e-3, weight_decay=wd, device=DEV)
    ref_p = [v.clone().to(DEV).requires_grad_(True) for v in params.values()]
    ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=wd, foreach=False, fused=False)
    for step in range(5):
        grads = [torch.randn(v.shape, generator=g) * (0.5 ** step) for v in params.values()]
        for k, gr, rp in zip(opt.names, grads, ref_p):
            opt.view(opt.grad, k).copy_(gr)
            rp.grad = gr.to(DEV)
        opt.step()
        ref.step()
        for k, rp in zip(opt.names, ref_p):
            assert (opt.view(opt.param, k) - rp.detach()).abs().max().item() <= 2e-7 * (1 + rp.detach().abs().max().item())
    # padding between the 16-byte aligned segments never moves
    assert opt.numel % 4 == 0 and opt.offsets["b"][0] % 4 == 0
    # the reference's own first Adam step (lr 1e-4, weight_decay 0): parameter deltas of the golden gradients
    if wd == 0.0:
        gd = _golden(golden_dir)
        names = [k[len("grad::"):] for k in gd.files if k.startswith("grad::")]
        usd = synth.synth_unet_state(seed=0)           # the parameters the reference started from
        before = {k: usd[k].clone() for k in names}
        o2 = train.FlatAdam(before, lr=1e-4, device=DEV)
        for k in names:
            o2.view(o2.grad, k).copy_(torch.from_numpy(gd[f"grad::{k}"]))
        o2.step()
        for k in names:
            delta = o2.view(o2.param, k).cpu() - before[k]
            assert (delta - torch.from_numpy(gd[f"delta::{k}"])).abs().max().item() <= 2e-8


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("N,co,ci,H,W", [(3, 128, 128, 16, 16), (2, 64, 17, 32, 32), (5, 256, 192, 8, 8), (9, 192, 320, 4, 4), (11, 128, 64, 2, 2)])
def test_conv_wgrad(N, co, ci, H, W, split):
    """dW of a 3x3 zero-padded conv against torch's conv2d weight gradient; split: bf16 hi + lo operands (fp32-class), else
    IEEE fp16 operands.  Ragged channel counts (17, 192, 320), small maps with several images per pixel box, N not a multiple
    of the box."""
    no_tf32()
    g = torch.Generator().manual_seed(N + co + ci)
    x = (torch.randn(N, ci, 1, H, W, generator=g)).to(DEV)
    dy = (torch.randn(N, co, 1, H, W, generator=g) * 0.3).to(DEV)
    if not split:
        x, dy = f16_round(x), f16_round(dy)
    xa, dya = to_act(x, split=split, f16=not split), to_act(dy, split=split, f16=not split)
    dw = torch.zeros(co, ci + 5, 3, 3, device=DEV)                       # a 5-channel block in front: cin_off
    train.conv_wgrad(dya, xa, dw, co, ci, 5, _s())
    ref = torch.nn.grad.conv2d_weight(x[:, :, 0], (co, ci, 3, 3), dy[:, :, 0], padding=1)
    assert dw[:, :5].abs().max().item() == 0
    assert rel_l2(dw[:, 5:], ref) <= (2e-5 if split else 1e-5), rel_l2(dw[:, 5:], ref)
    # the trainer's layout [Cout, 3, 3, Cin]: 16-byte vector reductions where a thread's 32 columns are aligned
    for off in (0, 4, 5):
        dwc = torch.zeros(co, 3, 3, ci + off, device=DEV)
        train.conv_wgrad(dya, xa, dwc, co, ci, off, _s(), channels_last=True)
        assert dwc[..., :off].abs().max().item() == 0 if off else True
        assert rel_l2(dwc[..., off:].permute(0, 3, 1, 2), ref) <= (2e-5 if split else 1e-5)


@pytest.mark.parametrize("N,co,ci,H,W", [(2, 768, 256, 8, 8), (3, 64, 128, 1, 1), (2, 1024, 1024, 2, 2), (4, 200, 72, 4, 8)])
def test_linear_wgrad(N, co, ci, H, W):
    """dW of a 1x1 conv / Linear over tokens (in_proj, out_proj . proj_out): dW[o][i] = sum_p dY[p][o] X[p][i]."""
    no_tf32()
    g = torch.Generator().manual_seed(N + co + ci)
    x = torch.randn(N, ci, 1, H, W, generator=g).to(DEV)
    dy = (torch.randn(N, co, 1, H, W, generator=g) * 0.3).to(DEV)
    dw = torch.zeros(co, ci, device=DEV)
    train.conv_wgrad(to_act(dy, split=True), to_act(x, split=True), dw, co, ci, 0, _s(), kind=train.WGRAD_LINEAR)
    ref = torch.einsum("nohw,nihw->oi", dy[:, :, 0].double(), x[:, :, 0].double()).float()
    assert rel_l2(dw, ref) <= 2e-5, rel_l2(dw, ref)


@pytest.mark.parametrize("N,ci,co,H,W", [(2, 128, 64, 16, 16), (3, 2048, 1024, 1, 1), (2, 256, 128, 4, 4), (5, 192, 72, 2, 8)])
def test_conv_transpose_gradients(N, ci, co, H, W):
    """ConvTranspose2d k2 s2 (unet/blocks.py:205): weight gradient [Cin, Cout, 2, 2] on the tcgen05 wgrad kernel (dY read
    with a TMA element stride of 2), bias gradient, and the data gradient as a stride-2 2x2 conv on the forward engine."""
    no_tf32()
    g = torch.Generator().manual_seed(N + co + ci)
    x = torch.randn(N, ci, H, W, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(ci, co, 2, 2, generator=g) / (ci ** 0.5)).to(DEV).requires_grad_(True)
    b = torch.randn(co, generator=g).to(DEV).requires_grad_(True)
    dy = (torch.randn(N, co, 2 * H, 2 * W, generator=g) * 0.3).to(DEV)
    F.conv_transpose2d(x, w, b, stride=2).backward(dy)
    xa, dya = to_act(x.detach()[:, :, None], split=True), to_act(dy[:, :, None], split=True)
    dw = torch.zeros(ci, co, 2, 2, device=DEV)
    train.conv_wgrad(dya, xa, dw, co, ci, 0, _s(), kind=train.WGRAD_CONVT2X2)
    assert rel_l2(dw, w.grad) <= 2e-5, rel_l2(dw, w.grad)
    dwc = torch.zeros(ci, 2, 2, co, device=DEV)
    train.conv_wgrad(dya, xa, dwc, co, ci, 0, _s(), kind=train.WGRAD_CONVT2X2, channels_last=True)
    assert rel_l2(dwc.permute(0, 3, 1, 2), w.grad) <= 2e-5
    db = torch.zeros(co, device=DEV)
    train.channel_sum(dya, db, co, _s())
    assert rel_l2(db, b.grad) <= 1e-5
    if co % 64 == 0:
        pd = train.pack_convT2x2_dgrad(w.detach().cpu(), DEV, split=True)
        dx = new_act(N, 1, H, W, xa.C, DEV, split=True)
        engine.ConvPlan([dya], pd, dx, cout=ci, stride=2).run(_s())
        assert rel_l2(from_act(dx, ci)[:, :, 0], x.grad) <= 2e-5, rel_l2(from_act(dx, ci)[:, :, 0], x.grad)


def test_add_and_maxpool_backward():
    g = torch.Generator().manual_seed(5)
    for N, C, H, W in ((2, 64, 32, 32), (3, 128, 4, 2), (1, 1024, 2, 2)):
        x = torch.randn(N, C, H, W, generator=g)
        x[:, :, 0, 0] = x[:, :, 0, 1]                       # ties: torch keeps the first maximum in scan order
        x[:, : C // 2, 1, 0] = x[:, : C // 2, 1, 1] = 9.0
        xa = to_act(x.to(DEV)[:, :, None], split=True)
        xr = from_act(xa, C)[:, :, 0].clone().requires_grad_(True)   # the values the kernel sees
        dy = torch.randn(N, C, H // 2, W // 2, generator=g).to(DEV)
        F.max_pool2d(xr, 2, 2).backward(dy)
        dya = to_act(dy[:, :, None], split=True)
        dxa = new_act(N, 1, H, W, C, DEV, split=True)
        train.maxpool_bwd(xa, dya, dxa, _s())
        assert torch.equal(from_act(dxa, C)[:, :, 0], from_act(dya, C)[:, :, 0].repeat_interleave(2, 2).repeat_interleave(2, 3) * (xr.grad != 0))
        assert rel_l2(from_act(dxa, C)[:, :, 0], xr.grad) <= 1e-5
        # skip connection: the two gradient paths add
        o = new_act(N, 1, H, W, C, DEV, split=True)
        train.add_acts(xa, dxa, o, _s())
        assert rel_l2(from_act(o, C), from_act(xa, C) + from_act(dxa, C)) <= 1e-5


@pytest.mark.parametrize("N,T,C,heads", [(2, 64, 256, 2), (3, 256, 128, 1), (2, 16, 512, 1), (2, 16, 1024, 8), (1, 64, 1024, 2), (3, 4, 1024, 2), (2, 40, 128, 4)])
def test_attention_core_backward(N, T, C, heads):
    """d(q | k | v) of softmax(q k^T / sqrt(d)) v against torch autograd (fp32)."""
    no_tf32()
    g = torch.Generator().manual_seed(T + C + heads)
    qkv = torch.randn(N, T, 3 * C, generator=g).to(DEV).requires_grad_(True)
    d_out = (torch.randn(N, T, C, generator=g) * 0.2).to(DEV)
    d = C // heads
    q, k, v = [t.reshape(N, T, heads, d).transpose(1, 2) for t in qkv.split(C, dim=2)]
    out = (torch.softmax(q @ k.transpose(-1, -2) / d ** 0.5, dim=-1) @ v).transpose(1, 2).reshape(N, T, C)
    out.backward(d_out)
    cl = lambda t: to_act(t.detach().transpose(1, 2).reshape(N, -1, 1, 1, T), split=True)
    qa, oa, da = cl(qkv), cl(out), cl(d_out)
    dq = new_act(N, 1, 1, T, 3 * C, DEV, split=True)
    train.attention_bwd(qa, oa, da, dq, heads, _s())
    got = from_act(dq, 3 * C).reshape(N, 3 * C, T).transpose(1, 2)
    for i, name in enumerate("qkv"):
        e = rel_l2(got[..., i * C:(i + 1) * C], qkv.grad[..., i * C:(i + 1) * C])
        assert e <= 2e-5, (name, e)


@pytest.mark.parametrize("N,segs,cmid,cout,H", [(3, [128], 128, 128, 16), (2, [64, 64], 64, 64, 32), (8, [256], 512, 512, 4), (2, [17], 64, 64, 32)])
def test_double_block_backward_vs_autograd(N, segs, cmid, cout, H):
    """One DoubleBlock (unet/blocks.py:50-107) forward + backward on the GPU against torch.autograd through the oracle's
    functional block: encoder block, decoder block over a channel concatenation, a small-map level, and the UNet's first
    block (17 input channels, no input gradient)."""
    no_tf32()
    g = torch.Generator().manual_seed(N + cmid + H)
    cin = sum(segs)
    sd = {"b.block1.conv.weight": torch.randn(cmid, cin, 3, 3, generator=g) * (9 * cin) ** -0.5,
          "b.block1.norm.weight": 1 + 0.2 * torch.randn(cmid, generator=g), "b.block1.norm.bias": 0.2 * torch.randn(cmid, generator=g),
          "b.block2.conv.weight": torch.randn(cout, cmid, 3, 3, generator=g) * (9 * cmid) ** -0.5,
          "b.block2.norm.weight": 1 + 0.2 * torch.randn(cout, generator=g), "b.block2.norm.bias": 0.2 * torch.randn(cout, generator=g)}
    x = torch.randn(N, cin, H, H, generator=g)
    tc = torch.randn(N, cmid, generator=g) * 0.5
    d_out = torch.randn(N, cout, H, H, generator=g)
    # reference: autograd through the oracle block (time embedding entering as the already-projected per-channel vector)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr, tr = x.clone().requires_grad_(True), tc.clone().requires_grad_(True)
    h = ounet._block(P, "b.block1", xr) + tr[:, :, None, None]
    out_ref = ounet._block(P, "b.block2", h)
    out_ref.backward(d_out)
    blk = train.DoubleBlockGrad(sd["b.block1.conv.weight"], sd["b.block1.norm.weight"], sd["b.block1.norm.bias"],
                                sd["b.block2.conv.weight"], sd["b.block2.norm.weight"], sd["b.block2.norm.bias"], segs, device=DEV)
    ins, c0 = [], 0
    for cs in segs:
        ins.append(to_act(x[:, c0:c0 + cs, None].to(DEV), split=True))
        c0 += cs
    out = blk.forward(ins, tc)
    assert rel_err(from_act(out, cout)[:, :, 0].cpu(), out_ref.detach()) <= 1e-3
    gr = blk.backward(to_act(d_out[:, :, None].to(DEV), split=True))
    want = {"conv1.weight": P["b.block1.conv.weight"].grad, "norm1.weight": P["b.block1.norm.weight"].grad, "norm1.bias": P["b.block1.norm.bias"].grad,
            "conv2.weight": P["b.block2.conv.weight"].grad, "norm2.weight": P["b.block2.norm.weight"].grad, "norm2.bias": P["b.block2.norm.bias"].grad,
            "temb": tr.grad}
    for k, ref in want.items():
        e = rel_l2(gr[k].cpu(), ref)
        assert e <= 2e-3, (k, e)
        assert e <= 2e-4, (k, e)   # the measured level of the fp32-class mode, an order inside the bound
    c0 = 0
    for cs, dx in zip(segs, gr["inputs"]):
        if cs % 64 == 0:
            e = rel_l2(from_act(dx, cs)[:, :, 0].cpu(), xr.grad[:, c0:c0 + cs])
            assert e <= 2e-4, e
        else:
            assert dx is None
        c0 += cs


def _train_inputs(golden_dir):
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("make_train_golden", os.path.join(golden_dir, "make_train_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    saved = list(sys.path)
    try:
        spec.loader.exec_module(mod)   # train_inputs() / FULL only: the reference is not imported at module level
    finally:
        sys.path[:] = saved
    return mod


def test_unet_training_step_vs_reference(golden_dir):
    """The whole training step (q_sample -> UNet -> criterion -> backward -> Adam) on the GPU against (a) the golden vectors
    of the unmodified reference: loss, eps-prediction, the norm of every one of the 172 parameter gradients, nine full
    gradient tensors and their post-Adam parameter deltas; (b) every gradient tensor of the oracle's autograd restatement
    (itself pinned to the same golden file by tests/test_oracle_golden.py)."""
    no_tf32()
    mod = _train_inputs(golden_dir)
    g = _golden(golden_dir)
    sd = synth.synth_unet_state(seed=0)
    x_start, cond, feats, noise, t = mod.train_inputs()
    tr = train.UNetTrainer(sd, **synth.UNET_KWARGS, lr=1e-4, weight_decay=0.0, device=DEV)
    before = {k: v.clone() for k, v in tr.state_dict().items()}
    loss, pred = tr.training_step(x_start, cond, feats, t, noise)
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    assert rel_err(pred.cpu(), torch.from_numpy(g["pred"])) <= 1e-4
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(tr.opt.names)
    worst = 0.0
    for n, ref_norm in zip(names, g["grad_norms"]):
        got = float(tr.G(n).double().norm())
        worst = max(worst, abs(got - ref_norm) / max(ref_norm, 1e-6))
        assert abs(got - ref_norm) <= 2e-3 * max(ref_norm, 1e-6), (n, got, ref_norm)
    for k in mod.FULL:
        ref_g = torch.from_numpy(g[f"grad::{k}"])
        assert rel_l2(tr.G(k).cpu(), ref_g) <= 2e-3, (k, rel_l2(tr.G(k).cpu(), ref_g))
        ref_d = torch.from_numpy(g[f"delta::{k}"])
        got_d = (tr.state_dict()[k] - before[k]).cpu()
        # Adam's first step moves every weight by ~lr * sign(g); entries whose gradient is ~0 flip on rounding noise
        big = ref_g.abs() > 1e-3 * ref_g.abs().max()
        assert (got_d - ref_d)[big].abs().max().item() <= 2e-6, k
    _, ograds, _ = otrain.training_loss_and_grads(sd, x_start, cond, feats, t, noise)
    for n in names:
        e = rel_l2(tr.G(n).cpu(), ograds[n])
        assert e <= 2e-3, (n, e)
    print(f"training step: loss {loss.item():.6f}, worst gradient-norm deviation {worst:.2e}")


def test_operand_refresh_in_place():
    """b2d_pack_weight rewrites every operand form of the step (forward / data-gradient conv, linear and its transpose,
    transposed conv and its data gradient) bit-identically to a fresh engine.pack_* of the same parameters."""
    g = torch.Generator().manual_seed(11)
    s = _s()
    w = torch.randn(96, 64 + 17, 3, 3, generator=g).to(DEV)
    wl = torch.randn(200, 72, generator=g).to(DEV)
    wt = torch.randn(128, 72, 2, 2, generator=g).to(DEV)
    live = [train.live_conv2d(w, [64, 17], None, DEV), train.live_conv2d_dgrad(w, (0, 64), DEV), train.live_linear(wl, None, DEV),
            train.live_linear(wl, None, DEV, transpose=True), train.live_convT2x2(wt, None, DEV), train.live_convT2x2_dgrad(wt, DEV)]
    for t in (w, wl, wt):
        t.copy_(torch.randn(t.shape, generator=g).to(DEV))         # "optimizer step"
    fresh = [train.live_conv2d(w, [64, 17], None, DEV), train.live_conv2d_dgrad(w, (0, 64), DEV), train.live_linear(wl, None, DEV),
             train.live_linear(wl, None, DEV, transpose=True), train.live_convT2x2(wt, None, DEV), train.live_convT2x2_dgrad(wt, DEV)]
    for a, b in zip(live, fresh):
        assert not torch.equal(a.pw.w.view(torch.int16), b.pw.w.view(torch.int16))
        a.refresh(s)
        assert torch.equal(a.pw.w.view(torch.int16), b.pw.w.view(torch.int16))
    # the same operands from channels-last parameters ([Cout, 3, 3, Cin], [Cin, 2, 2, Cout]: the trainer's layout)
    wc, wtc = w.permute(0, 2, 3, 1).contiguous(), wt.permute(0, 2, 3, 1).contiguous()
    live_cl = [train.live_conv2d(wc, [64, 17], None, DEV, channels_last=True), train.live_conv2d_dgrad(wc, (0, 64), DEV, channels_last=True),
               train.live_convT2x2(wtc, None, DEV, channels_last=True), train.live_convT2x2_dgrad(wtc, DEV, channels_last=True)]
    for a, b in zip(live_cl, [fresh[0], fresh[1], fresh[4], fresh[5]]):
        assert torch.equal(a.pw.w.view(torch.int16), b.pw.w.view(torch.int16))
        a.pw.w.zero_()
        a.refresh(s)
        assert torch.equal(a.pw.w.view(torch.int16), b.pw.w.view(torch.int16))


def test_second_training_step_reuses_buffers_and_refreshed_operands(golden_dir):
    """Step 2 of a trainer (cached buffers and plans, operands rewritten in place after step 1's update) against a trainer
    built fresh from the step-1 parameters."""
    no_tf32()
    mod = _train_inputs(golden_dir)
    x_start, cond, feats, noise, t = mod.train_inputs()
    a = train.UNetTrainer(synth.synth_unet_state(seed=0), **synth.UNET_KWARGS, lr=1e-3, device=DEV)
    a.training_step(x_start, cond, feats, t, noise)
    b = train.UNetTrainer({k: v.clone() for k, v in a.state_dict().items()}, **synth.UNET_KWARGS, lr=1e-3, device=DEV)
    x = torch.cat([x_start, cond, feats], dim=1).to(DEV)
    la, pa = a.forward_backward(x, t, noise.to(DEV))
    lb, pb = b.forward_backward(x, t, noise.to(DEV))
    torch.cuda.synchronize()
    assert abs(la.item() - lb.item()) <= 1e-6 * abs(lb.item())
    assert torch.equal(pa, pb)                                       # the forward has no atomics: bit-identical
    for n in a.opt.names:
        assert rel_l2(a.G(n), b.G(n)) <= 1e-5, n                     # wgrad sums pixel splits with fp32 atomics
    # a different batch shape on the same trainer: new buffers and plans, same operands
    a.forward_backward(x[:1, :, :32, :32].contiguous(), t[:1], noise[:1].to(DEV))
    l3, _ = a.forward_backward(x, t, noise.to(DEV))
    assert abs(l3.item() - lb.item()) <= 1e-6 * abs(lb.item())


def test_bf16_training_step_tracks_the_reference(golden_dir):
    """The single-pass mode (bf16 operands / activations / gradients, fp32 accumulate): a third of the MMA work.  Measured
    against the oracle's fp32 autograd: loss to ~1e-3, the median parameter gradient to ~6e-2, the worst (GroupNorm weights) to ~0.12 (rel-L2) on the 2 x 32 x 32 golden step --
    a throughput mode, not the parity mode."""
    mod = _train_inputs(golden_dir)
    g = _golden(golden_dir)
    sd = synth.synth_unet_state(seed=0)
    x_start, cond, feats, noise, t = mod.train_inputs()
    tr = train.UNetTrainer(sd, **synth.UNET_KWARGS, device=DEV, precision="bf16")
    loss, pred = tr.training_step(x_start, cond, feats, t, noise)
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) <= 5e-3 * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    _, ograds, _ = otrain.training_loss_and_grads(sd, x_start, cond, feats, t, noise)
    errs = {n: rel_l2(tr.G(n).cpu(), ograds[n]) for n in tr.opt.names}
    worst = max(errs, key=errs.get)
    print(f"bf16 training step: loss {loss.item():.5f} (reference {float(g['loss']):.5f}), worst gradient rel-L2 {errs[worst]:.3e} ({worst}), "
          f"median {sorted(errs.values())[len(errs) // 2]:.3e}")
    # measured: worst 0.12 (a GroupNorm weight: its gradient is a cancelling sum over dy * xhat), median 5.8e-2
    assert errs[worst] <= 0.25, (worst, errs[worst])
    assert sorted(errs.values())[len(errs) // 2] <= 0.1
    with pytest.raises(ValueError):
        train.UNetTrainer(sd, **synth.UNET_KWARGS, device=DEV, precision="fp8")
    with pytest.raises(NotImplementedError):
        train.UNetTrainer(sd, **{**synth.UNET_KWARGS, "padding_mode": "reflect"}, device=DEV)


def test_training_step_at_baseline_size():
    """BASELINE configs[4]'s per-GPU step: 2 samples = 22 slice-images of 8 x 64 x 64 latents, graph-replayed (third call),
    every parameter gradient against the oracle's fp32 autograd on the host."""
    no_tf32()
    g = torch.Generator().manual_seed(77)
    N, S = 22, 64
    x_start, cond = torch.randn(N, 8, S, S, generator=g), torch.randn(N, 8, S, S, generator=g)
    feats, noise = torch.rand(N, 1, S, S, generator=g), torch.randn(N, 8, S, S, generator=g)
    t = torch.randint(0, 1000, (N,), generator=g)
    sd = synth.synth_unet_state(seed=0)
    tr = train.UNetTrainer(sd, **synth.UNET_KWARGS, lr=0.0, device=DEV)     # lr 0: three identical steps, the third is a replay
    for _ in range(3):
        loss, pred = tr.training_step(x_start, cond, feats, t, noise)
    torch.cuda.synchronize()
    assert tr._graph["fb"] is not None
    oloss, ograds, opred = otrain.training_loss_and_grads(sd, x_start, cond, feats, t, noise)
    assert abs(loss.item() - oloss.item()) <= 1e-5 * abs(oloss.item()), (loss.item(), oloss.item())
    assert rel_err(pred.cpu(), opred) <= 1e-4
    errs = {n: rel_l2(tr.G(n).cpu(), ograds[n]) for n in tr.opt.names}
    worst = max(errs, key=errs.get)
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print(f"training step at 22 x 64 x 64: loss {loss.item():.6f}, largest gradient rel-L2 errors {[(k, f'{v:.1e}') for k, v in top]}, "
          f"median {sorted(errs.values())[len(errs) // 2]:.1e}")
    # measured: median 8.7e-5; largest 3.2e-3 (encoder.1.2.norm.weight), 2.8e-3 (encoder.2.0.block1.conv.weight).  Those are the
    # tensors tests/diag_train_sensitivity.py finds most sensitive to the mode's 2^-17 storage rounding: quantising only
    # the max-pool inputs inside the oracle's own fp32 step moves them by 2.1e-3 / 1.7e-3 (three argmax flips in 2.8 M
    # windows), while the fp32 oracle is within 4.5e-6 of fp64.  Bound: 5e-3 on the worst tensor, 3e-4 on the median.
    assert errs[worst] <= 5e-3, (worst, errs[worst])
    assert sorted(errs.values())[len(errs) // 2] <= 3e-4


def _field_predictor(precision):
    from diffusion_model_project_b200.predictor import B200LatentDiffusionPredictor
    vsd = synth.synth_vae_state(seed=1, branches=("encoder_2d", "encoder_3d", "decoder_3d"))
    return B200LatentDiffusionPredictor("UNet", dict(synth.UNET_KWARGS), True, unet_state=synth.synth_unet_state(seed=0), vae_state=vsd,
                                        norm_factors=synth.NORM_FACTORS, num_slices=2, num_timesteps=1000, precision=precision, device=DEV)


@pytest.mark.parametrize("precision,bound", [("fp32x", 1e-3), ("f16", 2e-2)])
def test_predictor_forward_vs_reference(golden_dir, precision, bound):
    """`predictor(img, velocity_2d, x_start=latents, noise=noise)` (predictor.py:636-751: frozen E2D conditioning + EDT
    features, q_sample at the drawn timesteps, UNet) against the unmodified reference's own forward() on the same fields,
    target latents, noise and timesteps (tests/golden/train_from_fields.npz).  Bound: BASELINE's per-step noise-prediction
    tolerance of the mode (max relative error)."""
    no_tf32()
    mod = _train_inputs(golden_dir)
    g = np.load(os.path.join(golden_dir, "train_from_fields.npz"))
    img, v2d, target, noise, _ = mod.field_inputs()
    t = torch.from_numpy(g["t"])
    with torch.no_grad():
        p = _field_predictor(precision)
        lat = p.encode_target(target.to(DEV), v2d.to(DEV))
        ref_lat = torch.from_numpy(g["latents"])
        e_lat = rel_err(lat.cpu(), ref_lat)
        pred, nz = p(img.to(DEV), v2d.to(DEV), x_start=torch.from_numpy(g["latents"]).to(DEV), noise=noise.to(DEV), t=t)
        torch.cuda.synchronize()
        e = rel_err(pred.cpu(), torch.from_numpy(g["pred"]))
        print(f"predictor.forward [{precision}]: encode_target max-rel {e_lat:.2e}, noise prediction max-rel {e:.2e} (bound {bound:g})")
        assert e_lat <= (1e-3 if precision == "fp32x" else 1e-2)
        assert e <= bound
        assert torch.equal(nz.cpu(), noise.reshape(nz.shape))
        with pytest.raises(ValueError):
            p(img.to(DEV), v2d.to(DEV))  # predictor.py:750: forward() without x_start
        # without injection the timesteps and the noise are drawn per call, like the reference's randint / randn_like
        a, na = p(img.to(DEV), v2d.to(DEV), x_start=lat)
        b, nb = p(img.to(DEV), v2d.to(DEV), x_start=lat)
        assert not torch.equal(na, nb) and not torch.equal(a, b)


def test_training_step_from_fields_vs_reference(golden_dir):
    """The reference's training loop body from the fields (helper.py:277-430: encode_target -> predictor.forward ->
    criterion -> backward -> Adam) as `LatentDiffusionTrainer.train_step`: frozen E3D / E2D passes, EDT and bilinear
    features of the B200 predictor (fp32-class mode), then the UNetTrainer step -- against one iteration of the unmodified
    reference predictor + torch.optim.Adam on the same fields, timesteps and noise (tests/golden/train_from_fields.npz)."""
    no_tf32()
    mod = _train_inputs(golden_dir)
    g = np.load(os.path.join(golden_dir, "train_from_fields.npz"))
    img, v2d, target, noise, _ = mod.field_inputs()
    t = torch.from_numpy(g["t"])
    usd = synth.synth_unet_state(seed=0)
    with torch.no_grad():
        p = _field_predictor("fp32x")
        lt = train.LatentDiffusionTrainer(p, unet_state=usd, lr=1e-4, weight_decay=0.0)
        before = {k: v.clone() for k, v in lt.unet.state_dict().items()}
        loss, pred, nz = lt.train_step(img.to(DEV), v2d.to(DEV), target.to(DEV), t=t, noise=noise)
        torch.cuda.synchronize()
    e_loss = abs(loss.item() - float(g["loss"])) / abs(float(g["loss"]))
    e_pred = rel_err(pred.cpu(), torch.from_numpy(g["pred"]))
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(lt.unet.opt.names)
    worst, worst_n = 0.0, ""
    for n, ref_norm in zip(names, g["grad_norms"]):
        d = abs(float(lt.unet.G(n).double().norm()) - ref_norm) / max(ref_norm, 1e-6)
        if d > worst:
            worst, worst_n = d, n
    full = {}
    for k in ("final_conv.weight", "encoder.0.0.block1.conv.weight", "bottleneck.block2.norm.weight"):
        ref_g = torch.from_numpy(g[f"grad::{k}"])
        full[k] = rel_l2(lt.unet.G(k).cpu(), ref_g)
        ref_d = torch.from_numpy(g[f"delta::{k}"])
        got_d = (lt.unet.state_dict()[k] - before[k]).cpu()
        big = ref_g.abs() > 1e-2 * ref_g.abs().max()
        assert (got_d - ref_d)[big].abs().max().item() <= 5e-6, k
    print(f"training step from fields: loss {loss.item():.6f} vs {float(g['loss']):.6f} (rel {e_loss:.2e}), noise prediction max-rel {e_pred:.2e}, "
          f"worst gradient-norm deviation {worst:.2e} ({worst_n}), full gradients rel-L2 {', '.join(f'{v:.2e}' for v in full.values())}")
    # measured on a B200: loss 2.2e-7, noise prediction 5.5e-5, worst gradient norm 7.3e-5, full gradients <= 5.3e-4
    assert e_loss <= 1e-5 and e_pred <= 5e-4
    assert worst <= 2e-3, (worst_n, worst)
    assert max(full.values()) <= 5e-3, full
    # the optimised parameters go back into the predictor's sampling UNet
    lt.sync_predictor()
    with torch.no_grad():
        pred2, _ = p(img.to(DEV), v2d.to(DEV), x_start=torch.from_numpy(g["latents"]).to(DEV), noise=noise.to(DEV), t=t)
    assert torch.isfinite(pred2).all() and not torch.equal(pred2, pred)
