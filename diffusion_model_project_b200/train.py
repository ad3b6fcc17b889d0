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

    def __init__(self, w1, g1, b1, w2, g2, b2, seg_sizes: Sequence[int], device="cuda", workspace: Optional[torch.Tensor] = None):
        dev = torch.device(device)
        self.dev = dev
        self.seg_sizes = list(seg_sizes)
        self.w1, self.w2 = w1.detach().float(), w2.detach().float()
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
        self.ws = workspace if workspace is not None else engine.new_workspace(dev)
        self.saved = None

    def forward(self, inputs: Sequence[Act], temb: Optional[torch.Tensor] = None, stats_out: Optional[torch.Tensor] = None) -> Act:
        """stats_out: (N, 2) fp64, receives (sum, sumsq) of the block's output (the next layer's GroupNorm statistics)."""
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
        engine.gn_apply(raw2, out, st2, self.cout, self.g2, self.b2, True, s, stats_out=stats_out)
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


class AttentionGrad:
    """Forward + backward of SelfAttention2d (unet/blocks.py:170-235): y = x + proj_out(out_proj(softmax(q k^T / sqrt(d)) v)),
    (q | k | v) = in_proj(GroupNorm(1, C)(x)).  The two C x C output projections run folded (W = Wp Wo, b = Wp bo + bp) as in
    the sampling path; their separate gradients follow from the folded ones by C x C parameter-space products."""

    def __init__(self, c: int, heads: int, p: Dict[str, torch.Tensor], device="cuda", workspace: Optional[torch.Tensor] = None):
        dev = torch.device(device)
        f = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self.dev, self.c, self.heads = dev, c, heads
        self.g, self.b = f(p["norm.weight"]), f(p["norm.bias"])
        self.w_in, self.b_in = f(p["mha.in_proj_weight"]), f(p["mha.in_proj_bias"])
        self.wo, self.bo = f(p["mha.out_proj.weight"]), f(p["mha.out_proj.bias"])
        self.wp, self.bp = f(p["proj_out.weight"])[:, :, 0], f(p["proj_out.bias"])
        w_out = (self.wp.double() @ self.wo.double()).float()
        b_out = (self.wp.double() @ self.bo.double() + self.bp.double()).float()
        self.p_in = engine.pack_linear(self.w_in, self.b_in, dev, split=True)
        self.p_out = engine.pack_linear(w_out, b_out, dev, split=True)
        self.p_in_t = engine.pack_linear(self.w_in.t().contiguous(), None, dev, split=True)
        self.p_out_t = engine.pack_linear(w_out.t().contiguous(), None, dev, split=True)
        self.ws = workspace if workspace is not None else engine.new_workspace(dev)
        self.saved = None

    def forward(self, x: Act, st_x: torch.Tensor) -> Act:
        """st_x: (N, 2) fp64 (sum, sumsq) of x per sample."""
        N, D, H, W, c = x.shape
        dev, s = self.dev, _lib.stream_ptr()
        xn = new_act(N, 1, H, W, c, dev, split=True)
        engine.gn_apply(x, xn, st_x, c, self.g, self.b, False, s)
        qkv = new_act(N, 1, H, W, 3 * c, dev, split=True)
        ConvPlan([xn], self.p_in, qkv, cout=3 * c, workspace=self.ws).run(s)
        ao = new_act(N, 1, H, W, c, dev, split=True)
        call("b2d_attention", ptr(qkv.hi), ptr(qkv.lo), ptr(ao.hi), ptr(ao.lo), N, H * W, c, self.heads, 0, s)
        y = new_act(N, 1, H, W, c, dev, split=True)
        ConvPlan([ao], self.p_out, y, cout=c, residual=x, workspace=self.ws).run(s)
        self.saved = (x, st_x, xn, qkv, ao)
        return y

    def backward(self, d_y: Act) -> Tuple[dict, Act]:
        x, st_x, xn, qkv, ao = self.saved
        N, D, H, W, c = x.shape
        dev, s = self.dev, _lib.stream_ptr()
        z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        d_wout, d_bout = z(c, c), z(c)
        conv_wgrad(d_y, ao, d_wout, c, c, 0, s, kind=WGRAD_LINEAR)
        channel_sum(d_y, d_bout, c, s)
        d_ao = new_act(N, 1, H, W, c, dev, split=True)
        ConvPlan([d_y], self.p_out_t, d_ao, cout=c, workspace=self.ws).run(s)
        d_qkv = new_act(N, 1, H, W, 3 * c, dev, split=True)
        attention_bwd(qkv, ao, d_ao, d_qkv, self.heads, s)
        g = {"mha.in_proj_weight": z(3 * c, c), "mha.in_proj_bias": z(3 * c), "norm.weight": z(c), "norm.bias": z(c)}
        conv_wgrad(d_qkv, xn, g["mha.in_proj_weight"], 3 * c, c, 0, s, kind=WGRAD_LINEAR)
        channel_sum(d_qkv, g["mha.in_proj_bias"], 3 * c, s)
        d_xn = new_act(N, 1, H, W, c, dev, split=True)
        ConvPlan([d_qkv], self.p_in_t, d_xn, cout=c, workspace=self.ws).run(s)
        d_x = new_act(N, 1, H, W, c, dev, split=True)
        gn_silu_bwd(x, d_xn, d_x, st_x, self.g, self.b, False, g["norm.weight"], g["norm.bias"], None, s)
        add_acts(d_x, d_y, d_x, s)
        # W = Wp Wo, b = Wp bo + bp  ->  dWp = dW Wo^T + db bo^T, dWo = Wp^T dW, dbo = Wp^T db, dbp = db
        dW, db = d_wout.double(), d_bout.double()
        g["proj_out.weight"] = (dW @ self.wo.double().t() + torch.outer(db, self.bo.double())).float()[:, :, None]
        g["proj_out.bias"] = d_bout
        g["mha.out_proj.weight"] = (self.wp.double().t() @ dW).float()
        g["mha.out_proj.bias"] = (self.wp.double().t() @ db).float()
        return g, d_x


class UNetTrainer:
    """One optimisation step of the conditioned eps-prediction UNet (helper.py:420-431, predictor.py:722-748):

        x_t = q_sample(x_start, t, noise);  pred = UNet(cat[x_t, cond, feats], t);  loss = criterion(pred, noise)
        loss.backward();  optimizer.step()                                     (torch.optim.Adam, train.py:144-148)

    in the fp32-class mode (bf16 hi + lo operands, three MMA passes): every layer's forward saves what its backward needs,
    the backward walks the UNet (unet/models.py:131-188) in reverse -- final_conv, decoder levels (attention, DoubleBlock over
    cat[skip, up], ConvTranspose2d + GroupNorm + SiLU), bottleneck, encoder levels (max-pool + GroupNorm + SiLU, attention,
    DoubleBlock) -- writing each parameter's gradient into FlatAdam's flat buffer, which one launch then applies.
    The sinusoid -> time_mlp -> per-block Linear chain ((N, 64) -> (N, 256) -> (N, Cmid): a few kFLOP) is evaluated and
    differentiated with torch ops on the GPU; everything that touches a feature map is libb2d."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], *, in_channels=17, out_channels=8, features=(64, 128, 256, 512, 1024),
                 attention: str = "", time_embedding_dim: Optional[int] = 64, num_timesteps: int = 1000, lr: float = 1e-4,
                 weight_decay: float = 0.0, device="cuda", **_ignored):
        from .scheduler import B200Scheduler
        from .synth import attention_heads
        if not torch.cuda.is_available():
            raise RuntimeError("UNetTrainer runs on a CUDA device only (no CPU fallback)")
        self.dev = torch.device(device)
        self.in_channels, self.out_channels, self.features = in_channels, out_channels, list(features)
        self.time_dim = time_embedding_dim
        self.heads = attention_heads(attention, len(self.features))
        self.opt = FlatAdam(state_dict, lr=lr, weight_decay=weight_decay, device=self.dev)
        self.scheduler = B200Scheduler(num_timesteps=num_timesteps, device=self.dev)
        self.ws = engine.new_workspace(self.dev)
        self._layers = None

    # -------------------------------------------------------------------------------- parameters
    def P(self, name: str) -> torch.Tensor:
        return self.opt.view(self.opt.param, name)

    def G(self, name: str) -> torch.Tensor:
        return self.opt.view(self.opt.grad, name)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return self.opt.state_dict()

    def _sub(self, prefix: str, names: Sequence[str]) -> Dict[str, torch.Tensor]:
        return {n: self.P(f"{prefix}.{n}") for n in names}

    _ATTN = ("norm.weight", "norm.bias", "mha.in_proj_weight", "mha.in_proj_bias", "mha.out_proj.weight", "mha.out_proj.bias",
             "proj_out.weight", "proj_out.bias")

    def _build_layers(self):
        """Operand forms of the current parameters (repacked after every optimizer step)."""
        dev, ws, f = self.dev, self.ws, self.features
        L = {}

        def double(p, segs):
            L[p] = DoubleBlockGrad(self.P(f"{p}.block1.conv.weight"), self.P(f"{p}.block1.norm.weight"), self.P(f"{p}.block1.norm.bias"),
                                   self.P(f"{p}.block2.conv.weight"), self.P(f"{p}.block2.norm.weight"), self.P(f"{p}.block2.norm.bias"),
                                   segs, dev, workspace=ws)

        cin = self.in_channels
        for lvl, c in enumerate(f):
            double(f"encoder.{lvl}.0", [cin])
            if self.heads[lvl] is not None:
                L[f"encoder.{lvl}.1"] = AttentionGrad(c, self.heads[lvl], self._sub(f"encoder.{lvl}.1", self._ATTN), dev, ws)
            cin = c
        double("bottleneck", [f[-1]])
        rheads = list(reversed(self.heads))
        for lvl, c in enumerate(reversed(f)):
            w = self.P(f"decoder.{lvl}.0.conv.weight")
            L[f"decoder.{lvl}.0"] = (engine.pack_convT2x2(w, self.P(f"decoder.{lvl}.0.conv.bias"), dev, split=True),
                                     pack_convT2x2_dgrad(w, dev, split=True))
            double(f"decoder.{lvl}.1", [c, c])
            if rheads[lvl] is not None:
                L[f"decoder.{lvl}.2"] = AttentionGrad(c, rheads[lvl], self._sub(f"decoder.{lvl}.2", self._ATTN), dev, ws)
        wf = self.P("final_conv.weight")
        L["final_conv"] = (engine.pack_conv2d(wf, [f[0]], self.P("final_conv.bias"), dev, split=True),
                           pack_conv2d_dgrad(wf, (0, f[0]), dev, split=True))
        self._layers = L

    # -------------------------------------------------------------------------------- time embedding chain (torch, tiny)
    def _time_names(self):
        nl = len(self.features)
        return [f"encoder.{l}.0" for l in range(nl)] + ["bottleneck"] + [f"decoder.{l}.1" for l in range(nl)]

    def _time_forward(self, t: torch.Tensor):
        import math
        import torch.nn.functional as F
        leaves = {}

        def leaf(name):
            leaves[name] = self.P(name).detach().clone().requires_grad_(True)
            return leaves[name]

        with torch.enable_grad():
            half = self.time_dim // 2
            fr = torch.exp(torch.arange(half, device=self.dev, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
            e = t.to(self.dev, torch.float32)[:, None] * fr[None, :]
            emb = torch.cat((e.sin(), e.cos()), dim=-1)                                  # models.py:14-26
            emb = F.linear(emb, leaf("time_mlp.0.weight"), leaf("time_mlp.0.bias"))      # models.py:78-82
            emb = F.linear(F.silu(emb), leaf("time_mlp.2.weight"), leaf("time_mlp.2.bias"))
            a = F.silu(emb)                                                              # blocks.py:92-95
            temb = {p: F.linear(a, leaf(f"{p}.time_mlp.1.weight"), leaf(f"{p}.time_mlp.1.bias")) for p in self._time_names()}
        return temb, leaves

    # -------------------------------------------------------------------------------- forward + backward
    def forward_backward(self, x: torch.Tensor, t: torch.Tensor, target: torch.Tensor):
        """x: (N, in_channels, h, w) fp32 on the GPU, t: (N,) timesteps, target: (N, out_channels, h, w).  Fills the flat
        gradient buffer; returns (loss, pred)."""
        if self._layers is None:
            self._build_layers()
        L, dev, f, s = self._layers, self.dev, self.features, _lib.stream_ptr()
        N, _, h, w = x.shape
        nl = len(f)
        if h % (1 << nl) or w % (1 << nl):
            raise ValueError(f"UNetTrainer: input {h}x{w} must be divisible by {1 << nl}")
        self.opt.zero_grad()
        z2 = lambda: torch.zeros(N, 2, dtype=torch.float64, device=dev)
        temb, leaves = self._time_forward(t) if self.time_dim is not None else ({}, {})
        x_in = new_act(N, 1, h, w, engine.pad64(self.in_channels), dev, split=True, zero=True)
        engine.planar_to_cl(x.contiguous().float(), x_in, N, self.in_channels, h * w, 0, None, s)
        # ---- forward (models.py:144-186)
        a, H, Wd = x_in, h, w
        skips, pools = [], []
        for lvl, c in enumerate(f):
            p = f"encoder.{lvl}.0"
            st = z2() if self.heads[lvl] is not None else None
            a = L[p].forward([a], temb.get(p).detach() if p in temb else None, stats_out=st)
            if st is not None:
                a = L[f"encoder.{lvl}.1"].forward(a, st)
            skips.append(a)
            praw = new_act(N, 1, H // 2, Wd // 2, c, dev, split=True)
            pst = z2()
            engine.maxpool_stats(a, praw, pst, s)
            pa = new_act(N, 1, H // 2, Wd // 2, c, dev, split=True)
            g, b = self.P(f"encoder.{lvl}.2.norm.weight"), self.P(f"encoder.{lvl}.2.norm.bias")
            engine.gn_apply(praw, pa, pst, c, g, b, True, s)
            pools.append((a, praw, pst))
            a, H, Wd = pa, H // 2, Wd // 2
        a = L["bottleneck"].forward([a], temb["bottleneck"].detach() if temb else None)
        ups = []
        rheads = list(reversed(self.heads))
        for lvl, c in enumerate(reversed(f)):
            pw, _ = L[f"decoder.{lvl}.0"]
            raw = new_act(N, 1, 2 * H, 2 * Wd, c, dev, split=True)
            ust = z2()
            ConvPlan([a], pw, raw, cout=c, nphase=4, stats=ust, stats_cpg=c, workspace=self.ws).run(s)
            up = new_act(N, 1, 2 * H, 2 * Wd, c, dev, split=True)
            engine.gn_apply(raw, up, ust, c, self.P(f"decoder.{lvl}.0.norm.weight"), self.P(f"decoder.{lvl}.0.norm.bias"), True, s)
            ups.append((a, raw, ust))
            H, Wd = 2 * H, 2 * Wd
            p = f"decoder.{lvl}.1"
            st = z2() if rheads[lvl] is not None else None
            a = L[p].forward([skips[nl - 1 - lvl], up], temb.get(p).detach() if p in temb else None, stats_out=st)
            if st is not None:
                a = L[f"decoder.{lvl}.2"].forward(a, st)
        pred = torch.empty(N, self.out_channels, h, w, dtype=torch.float32, device=dev)
        ConvPlan([a], L["final_conv"][0], pred, cout=self.out_channels, out_mode=1, out_cstride=self.out_channels, workspace=self.ws).run(s)
        final_in = a
        # ---- criterion (metrics.py:337-402)
        loss, _, d_pred = nmse_loss(pred, target.to(dev))
        # ---- backward
        d_temb = {}
        oc, c0 = self.out_channels, f[0]
        d = new_act(N, 1, h, w, engine.pad64(oc), dev, split=True, zero=True)
        engine.planar_to_cl(d_pred, d, N, oc, h * w, 0, None, s)
        channel_sum(d, self.G("final_conv.bias"), oc, s)
        conv_wgrad(d, final_in, self.G("final_conv.weight"), oc, c0, 0, s)
        da = new_act(N, 1, h, w, c0, dev, split=True)
        ConvPlan([d], L["final_conv"][1], da, cout=c0, workspace=self.ws).run(s)
        d_skips = [None] * nl

        def take_double(p, g):
            self.G(f"{p}.block1.conv.weight").copy_(g["conv1.weight"])
            self.G(f"{p}.block2.conv.weight").copy_(g["conv2.weight"])
            for i in (1, 2):
                self.G(f"{p}.block{i}.norm.weight").copy_(g[f"norm{i}.weight"])
                self.G(f"{p}.block{i}.norm.bias").copy_(g[f"norm{i}.bias"])
            if g["temb"] is not None:
                d_temb[p] = g["temb"]

        def take_attention(p, g):
            for k, v in g.items():
                self.G(f"{p}.{k}").copy_(v)

        for lvl in reversed(range(nl)):
            c = list(reversed(f))[lvl]
            if rheads[lvl] is not None:
                g, da = L[f"decoder.{lvl}.2"].backward(da)
                take_attention(f"decoder.{lvl}.2", g)
            g = L[f"decoder.{lvl}.1"].backward(da)
            take_double(f"decoder.{lvl}.1", g)
            d_skips[nl - 1 - lvl], d_up = g["inputs"]
            x_lo, raw, ust = ups[lvl]
            Nn, _, H2, W2, _ = raw.shape
            d_raw = new_act(N, 1, H2, W2, c, dev, split=True)
            gn_silu_bwd(raw, d_up, d_raw, ust, self.P(f"decoder.{lvl}.0.norm.weight"), self.P(f"decoder.{lvl}.0.norm.bias"), True,
                        self.G(f"decoder.{lvl}.0.norm.weight"), self.G(f"decoder.{lvl}.0.norm.bias"), None, s)
            conv_wgrad(d_raw, x_lo, self.G(f"decoder.{lvl}.0.conv.weight"), c, 2 * c, 0, s, kind=WGRAD_CONVT2X2)
            channel_sum(d_raw, self.G(f"decoder.{lvl}.0.conv.bias"), c, s)
            da = new_act(N, 1, H2 // 2, W2 // 2, 2 * c, dev, split=True)
            ConvPlan([d_raw], L[f"decoder.{lvl}.0"][1], da, cout=2 * c, stride=2, workspace=self.ws).run(s)
        g = L["bottleneck"].backward(da)
        take_double("bottleneck", g)
        da = g["inputs"][0]
        for lvl in reversed(range(nl)):
            c = f[lvl]
            skip, praw, pst = pools[lvl]
            d_praw = new_act(*praw.shape, dev, split=True)
            gn_silu_bwd(praw, da, d_praw, pst, self.P(f"encoder.{lvl}.2.norm.weight"), self.P(f"encoder.{lvl}.2.norm.bias"), True,
                        self.G(f"encoder.{lvl}.2.norm.weight"), self.G(f"encoder.{lvl}.2.norm.bias"), None, s)
            da = new_act(*skip.shape, dev, split=True)
            maxpool_bwd(skip, d_praw, da, s)
            add_acts(da, d_skips[lvl], da, s)
            if self.heads[lvl] is not None:
                g, da = L[f"encoder.{lvl}.1"].backward(da)
                take_attention(f"encoder.{lvl}.1", g)
            g = L[f"encoder.{lvl}.0"].backward(da)
            take_double(f"encoder.{lvl}.0", g)
            da = g["inputs"][0]
        if temb:
            names = list(d_temb.keys())
            torch.autograd.backward([temb[p] for p in names], [d_temb[p] for p in names])
            for k, v in leaves.items():
                self.G(k).copy_(v.grad)
        return loss, pred

    def training_step(self, x_start: torch.Tensor, cond: torch.Tensor, feats: torch.Tensor, t: torch.Tensor, noise: torch.Tensor,
                      group=None):
        """q_sample -> forward -> loss -> backward -> (gradient all-reduce) -> Adam.  Returns (loss, pred)."""
        dev = self.dev
        x_t = self.scheduler.q_sample(x_start.to(dev), t.to(dev), noise.to(dev))            # diffusion.py:78-101
        x = torch.cat([x_t, cond.to(dev), feats.to(dev)], dim=1)                             # predictor.py:731-741
        loss, pred = self.forward_backward(x, t, noise.to(dev))
        scale = self.opt.allreduce_gradients(group)
        self.opt.step(grad_scale=scale)
        self._layers = None  # operands are stale now
        return loss, pred
