"""First slice of the UNet training step over libb2d:

    q_sample(target latents) -> concat -> UNet eps-prediction -> normalized_mse_loss_per_component -> backward -> Adam
    (Diffusion_model/src/predictor.py:722-748, unet/metrics.py:337-402, helper.py:428-430, train.py:144-148)

What exists: the criterion (forward + gradient), torch.optim.Adam's update over flat fp32 parameter / moment buffers with
the gradient all-reduce that precedes it, and the backward of one DoubleBlock (unet/blocks.py:50-107: conv3x3 -> GroupNorm(1,C)
-> SiLU (+ time embedding) -> conv3x3 -> GroupNorm -> SiLU): GroupNorm/SiLU backward, the 3x3 conv's data gradient (the forward
engine on dY with mirrored taps and transposed weights) and its weight gradient (a tcgen05 kernel reading both operands
MN-major out of the channels-last tensors).  What does not exist yet: attention / max-pool / transposed-conv backward and
the time-MLP chain, i.e. the whole-UNet backward.  Gradients are carried in the fp32-class format (bf16 hi + lo): IEEE
fp16 would underflow them without loss scaling.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, engine
from ._lib import call, ptr
from .engine import Act, ConvPlan, new_act


# ------------------------------------------------------------------------------------------------ criterion
def nmse_loss(pred: torch.Tensor, target: torch.Tensor, weight_per_channel: Optional[torch.Tensor] = None, eps: float = 1e-8,
              want_grad: bool = True):
    """normalized_mse_loss_per_component (metrics.py:337-402), reduce=True.  pred, target: (N, C, *spatial) fp32 on the GPU.
    Returns (loss 0-d tensor, per-sample losses (N,), d loss / d pred or None)."""
    if pred.dim() not in (4, 5):
        raise ValueError(f"Expected 4D or 5D tensor, got {pred.dim()}D")  # metrics.py:369
    if pred.shape != target.shape:
        raise ValueError(f"shape mismatch: {tuple(pred.shape)} vs {tuple(target.shape)}")
    if not pred.is_cuda:
        raise RuntimeError("nmse_loss runs on a CUDA device only (no CPU fallback)")
    N, C = pred.shape[:2]
    P = pred[0, 0].numel()
    pred, target = pred.contiguous().float(), target.contiguous().float()
    err = torch.empty(N * C, dtype=torch.float32, device=pred.device)
    loss = torch.empty(1 + N, dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    w = None if weight_per_channel is None else weight_per_channel.to(pred.device, torch.float32).reshape(-1).contiguous()
    call("b2d_nmse_loss", pred.data_ptr(), target.data_ptr(), N, C, P, ptr(w), float(eps), err.data_ptr(), loss.data_ptr(), ptr(grad),
         _lib.stream_ptr(), launches=2)
    return loss[0], loss[1:], grad


# ------------------------------------------------------------------------------------------------ optimizer
class FlatAdam:
    """torch.optim.Adam (train.py:144-148: lr, weight_decay, default betas / eps) over ONE flat fp32 buffer holding every
    parameter back to back (16-byte aligned segments), with flat gradient and moment buffers of the same layout: one
    HBM-bound launch per step instead of one per tensor, and one contiguous buffer for the gradient all-reduce."""

    def __init__(self, params: Dict[str, torch.Tensor], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, device="cuda"):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.names: List[str] = list(params.keys())
        self.offsets: Dict[str, Tuple[int, int, torch.Size]] = {}
        off = 0
        for k in self.names:
            n = params[k].numel()
            self.offsets[k] = (off, n, params[k].shape)
            off += (n + 3) // 4 * 4
        self.numel = off
        dev = torch.device(device)
        self.param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        for k in self.names:
            self.view(self.param, k).copy_(params[k].detach().to(dev, torch.float32))
        self.step_count = 0

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        off, n, shape = self.offsets[name]
        return flat[off:off + n].view(shape)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self.view(self.param, k) for k in self.names}

    def zero_grad(self):
        self.grad.zero_()

    def allreduce_gradients(self, group=None) -> float:
        """Sum the flat gradient over the ranks (NCCL over NVLink on the GPU box; one collective for the whole model) and
        return the scale (1 / world) the update applies -- the mean of the reference's DataParallel-style replicas."""
        import torch.distributed as dist
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return 1.0
        dist.all_reduce(self.grad, group=group)
        return 1.0 / dist.get_world_size(group)

    def step(self, grad_scale: float = 1.0):
        if not self.param.is_cuda:
            raise RuntimeError("FlatAdam.step runs on a CUDA device only (no CPU fallback)")
        self.step_count += 1
        call("b2d_adam_step", self.param.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.numel,
             float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay), self.step_count,
             float(grad_scale), _lib.stream_ptr())


# ------------------------------------------------------------------------------------------------ layer backward
def pack_conv2d_dgrad(w: torch.Tensor, seg: Tuple[int, int], device, split=False, f16=False) -> engine.PackedWeight:
    """Data-gradient operand of nn.Conv2d 3x3 / pad 1 (weight [Cout, Cin, 3, 3]) for the input channels [seg0, seg1):
    dX[p][ci] = sum_{tap, co} dY[p - off(tap)][co] W[co][ci][tap]  ==  the forward engine on dY with rows = ci and the
    taps mirrored."""
    c0, c1 = seg
    return engine.pack_conv2d(w[:, c0:c1].transpose(0, 1).flip(2, 3).contiguous(), [w.shape[0]], None, device, split, f16)


def gn_silu_bwd(x: Act, dy: Act, dx: Act, stats: torch.Tensor, gamma, beta, act: bool, dgamma: torch.Tensor, dbeta: torch.Tensor,
                dtemb: Optional[torch.Tensor], stream: int, eps: float = 1e-5):
    N, D, H, W, C = x.shape
    sums = torch.empty(N, 2, dtype=torch.float64, device=x.hi.device)
    call("b2d_gn_silu_bwd", ptr(x.hi), ptr(x.lo), 1 if x.f16 else 0, ptr(dy.hi), ptr(dy.lo), 1 if dy.f16 else 0, ptr(dx.hi), ptr(dx.lo),
         1 if dx.f16 else 0, N, D * H * W, C, stats.data_ptr(), ptr(gamma), ptr(beta), float(eps), 1 if act else 0, sums.data_ptr(),
         dgamma.data_ptr(), dbeta.data_ptr(), ptr(dtemb), stream, launches=3)
    return sums


WGRAD_CONV3X3, WGRAD_LINEAR, WGRAD_CONVT2X2 = 0, 1, 2


def conv_wgrad(dy: Act, x: Act, dw: torch.Tensor, cout: int, cin: int, cin_off: int, stream: int, kind: int = WGRAD_CONV3X3):
    """dw (fp32, the reference's parameter layout) += the weight gradient of the input-channel block starting at cin_off.
    kind: WGRAD_CONV3X3 (dw [Cout, Cin_total, 3, 3]), WGRAD_LINEAR (dw [Cout, Cin_total]; Linear, Conv1d k1),
    WGRAD_CONVT2X2 (dw [Cin, Cout, 2, 2]; dy is the 2H x 2W gradient of the transposed conv's output)."""
    N, D, H, W, _ = x.shape
    up = 2 if kind == WGRAD_CONVT2X2 else 1
    assert D == 1 and dy.shape[:4] == (N, 1, H * up, W * up) and dw.dtype == torch.float32 and dw.is_contiguous()
    assert x.f16 == dy.f16 and (x.lo is None) == (dy.lo is None)
    cin_total = dw.shape[0] if kind == WGRAD_CONVT2X2 else dw.shape[1]
    call("b2d_conv_wgrad", kind, ptr(dy.hi), ptr(dy.lo), dy.C, ptr(x.hi), ptr(x.lo), x.C, N, H, W, cout, cin, cin_off, cin_total,
         dw.data_ptr(), 1 if x.f16 else 0, stream)


def channel_sum(x: Act, out: torch.Tensor, cvalid: int, stream: int):
    """out[c] += sum of x over every position (bias gradients)."""
    N, D, H, W, C = x.shape
    assert out.dtype == torch.float32 and out.numel() >= cvalid
    call("b2d_channel_sum", ptr(x.hi), ptr(x.lo), 1 if x.f16 else 0, N * D * H * W, C, cvalid, out.data_ptr(), stream)


def add_acts(a: Act, b: Act, out: Act, stream: int):
    """out = a + b (may alias either input)."""
    assert a.shape == b.shape == out.shape and a.f16 == b.f16 == out.f16
    call("b2d_add16", ptr(a.hi), ptr(a.lo), ptr(b.hi), ptr(b.lo), ptr(out.hi), ptr(out.lo), 1 if a.f16 else 0, a.hi.numel(), stream)


def maxpool_bwd(x: Act, dy: Act, dx: Act, stream: int):
    """MaxPool2d(2, 2) backward: x = the pooled layer's input, dy = gradient of its output."""
    N, D, H, W, C = x.shape
    assert D == 1 and dy.shape == (N, 1, H // 2, W // 2, C) and dx.shape == x.shape
    call("b2d_maxpool2x2_bwd", ptr(x.hi), ptr(x.lo), ptr(dy.hi), ptr(dy.lo), ptr(dx.hi), ptr(dx.lo), 1 if x.f16 else 0, N, H, W, C, stream)


def attention_bwd(qkv: Act, out: Act, d_out: Act, d_qkv: Act, heads: int, stream: int):
    """Backward of engine's attention core (b2d_attention): d_qkv = (dq | dk | dv)."""
    N, D, H, W, C3 = qkv.shape
    C, T = C3 // 3, D * H * W
    assert out.shape == d_out.shape == (N, D, H, W, C) and d_qkv.shape == qkv.shape
    stats = torch.empty(N * heads * T * 2, dtype=torch.float32, device=qkv.hi.device)
    call("b2d_attention_bwd", ptr(qkv.hi), ptr(qkv.lo), ptr(out.hi), ptr(out.lo), ptr(d_out.hi), ptr(d_out.lo), ptr(d_qkv.hi), ptr(d_qkv.lo),
         stats.data_ptr(), N, T, C, heads, 1 if qkv.f16 else 0, stream, launches=3)


def pack_convT2x2_dgrad(w: torch.Tensor, device, split=False, f16=False) -> engine.PackedWeight:
    """Data gradient of nn.ConvTranspose2d k2 s2 (weight [Cin, Cout, 2, 2]): dX[p][ci] = sum_{ky,kx,co} dY[2p + (ky,kx)][co]
    W[ci][co][ky][kx] -- a stride-2 conv with a 2x2 window (ConvPlan(..., stride=2))."""
    ci, co = w.shape[:2]
    taps = [(0, ky, kx) for ky in range(2) for kx in range(2)]
    return engine.pack_weight(w.permute(0, 2, 3, 1).reshape(ci, 4, co), [co], taps, None, device, split, f16=f16)


class DoubleBlockGrad:
    """Forward + backward of one DoubleBlock (unet/blocks.py:50-107) in the fp32-class mode, from the reference's own
    parameter tensors: conv{1,2}.weight [C, Cin, 3, 3] (no bias), norm{1,2}.{weight, bias}; `temb` (N, Cmid) is the per-sample
    time embedding added after block1 (blocks.py:100-103).  Inputs may be a channel concatenation (decoder blocks)."""

    def __init__(self, w1, g1, b1, w2, g2, b2, seg_sizes: Sequence[int], device="cuda"):
        dev = torch.device(device)
        self.dev = dev
        self.seg_sizes = list(seg_sizes)
        self.w1, self.w2 = w1.detach().float().cpu(), w2.detach().float().cpu()
        self.cmid, self.cout = w1.shape[0], w2.shape[0]
        f = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self.g1, self.b1, self.g2, self.b2 = f(g1), f(b1), f(g2), f(b2)
        self.pw1 = engine.pack_conv2d(self.w1, self.seg_sizes, None, dev, split=True)
        self.pw2 = engine.pack_conv2d(self.w2, [self.cmid], None, dev, split=True)
        self.pd2 = pack_conv2d_dgrad(self.w2, (0, self.cmid), dev, split=True)
        self.pd1, c0 = [], 0
        for cs in self.seg_sizes:
            # the data gradient of a segment is only defined for engine-sized channel counts (the UNet's first layer has
            # 17 input channels and needs none: its input is data)
            self.pd1.append(pack_conv2d_dgrad(self.w1, (c0, c0 + cs), dev, split=True) if cs % 64 == 0 else None)
            c0 += cs
        self.ws = engine.new_workspace(dev)
        self.saved = None

    def forward(self, inputs: Sequence[Act], temb: Optional[torch.Tensor] = None) -> Act:
        N, D, H, W, _ = inputs[0].shape
        dev, s = self.dev, _lib.stream_ptr()
        st1 = torch.zeros(N, 2, dtype=torch.float64, device=dev)
        st2 = torch.zeros(N, 2, dtype=torch.float64, device=dev)
        raw1 = new_act(N, 1, H, W, self.cmid, dev, split=True)
        a1 = new_act(N, 1, H, W, self.cmid, dev, split=True)
        raw2 = new_act(N, 1, H, W, self.cout, dev, split=True)
        out = new_act(N, 1, H, W, self.cout, dev, split=True)
        ConvPlan(list(inputs), self.pw1, raw1, cout=self.cmid, stats=st1, stats_cpg=self.cmid, workspace=self.ws).run(s)
        row = torch.arange(N, dtype=torch.int32, device=dev) if temb is not None else None
        tt = None if temb is None else temb.to(dev, torch.float32).contiguous()
        engine.gn_apply(raw1, a1, st1, self.cmid, self.g1, self.b1, True, s, temb=tt, temb_row=row, temb_row_stride=1)
        ConvPlan([a1], self.pw2, raw2, cout=self.cout, stats=st2, stats_cpg=self.cout, workspace=self.ws).run(s)
        engine.gn_apply(raw2, out, st2, self.cout, self.g2, self.b2, True, s)
        self.saved = (list(inputs), raw1, st1, a1, raw2, st2, temb is not None)
        return out

    def backward(self, d_out: Act) -> dict:
        """Returns {'conv1.weight', 'norm1.weight', 'norm1.bias', 'conv2.weight', 'norm2.weight', 'norm2.bias', 'temb' (N, Cmid),
        'inputs': [Act or None per segment]} -- the gradients torch.autograd gives for the same block."""
        inputs, raw1, st1, a1, raw2, st2, has_temb = self.saved
        N, D, H, W, _ = raw1.shape
        dev, s = self.dev, _lib.stream_ptr()
        z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        g = {"norm2.weight": z(self.cout), "norm2.bias": z(self.cout), "norm1.weight": z(self.cmid), "norm1.bias": z(self.cmid),
             "conv2.weight": z(self.cout, self.cmid, 3, 3), "conv1.weight": z(self.cmid, sum(self.seg_sizes), 3, 3),
             "temb": z(N, self.cmid) if has_temb else None}
        d_raw2 = new_act(N, 1, H, W, self.cout, dev, split=True)
        gn_silu_bwd(raw2, d_out, d_raw2, st2, self.g2, self.b2, True, g["norm2.weight"], g["norm2.bias"], None, s)
        conv_wgrad(d_raw2, a1, g["conv2.weight"], self.cout, self.cmid, 0, s)
        d_a1 = new_act(N, 1, H, W, self.cmid, dev, split=True)
        ConvPlan([d_raw2], self.pd2, d_a1, cout=self.cmid, workspace=self.ws).run(s)
        d_raw1 = new_act(N, 1, H, W, self.cmid, dev, split=True)
        gn_silu_bwd(raw1, d_a1, d_raw1, st1, self.g1, self.b1, True, g["norm1.weight"], g["norm1.bias"], g["temb"], s)
        d_inputs, c0 = [], 0
        for x, cs, pd in zip(inputs, self.seg_sizes, self.pd1):
            conv_wgrad(d_raw1, x, g["conv1.weight"], self.cmid, cs, c0, s)
            if pd is None:
                d_inputs.append(None)
            else:
                dx = new_act(N, 1, H, W, x.C, dev, split=True)
                ConvPlan([d_raw1], pd, dx, cout=cs, workspace=self.ws).run(s)
                d_inputs.append(dx)
            c0 += cs
        g["inputs"] = d_inputs
        return g
