"""The UNet training step over libb2d:

    q_sample(target latents) -> concat -> UNet eps-prediction -> normalized_mse_loss_per_component -> backward -> Adam
    (Diffusion_model/src/predictor.py:722-748, unet/metrics.py:337-402, helper.py:428-430, train.py:144-148)

* `nmse_loss`: the criterion, forward + gradient.
* `FlatAdam`: torch.optim.Adam's update over flat fp32 parameter / moment buffers, with the gradient all-reduce (whole, or
  by backward-order buckets) that precedes it.
* `DoubleBlockGrad`, `AttentionGrad` and the helpers above them: forward with saved activations and backward of the UNet's
  blocks -- GroupNorm / SiLU backward, conv / ConvTranspose2d / Linear data gradients (the forward engine on dY with mirrored
  taps and transposed weights), weight gradients (a tcgen05 kernel reading both operands MN-major out of the channels-last
  tensors), max-pool and attention-core backward.
* `UNetTrainer`: one whole optimisation step from latents, replayed as CUDA graphs, operands rewritten in place.
* `LatentDiffusionTrainer`: the reference's training loop body from the FIELDS (helper.py:277-430): frozen-VAE target
  latents and conditioning from a `B200LatentDiffusionPredictor`, then `UNetTrainer.training_step`.
Gradients are carried in the fp32-class format (bf16 hi + lo): IEEE fp16 would underflow them without loss scaling.
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

    def backward_order_buckets(self, boundaries: Sequence[str]) -> List[Tuple[int, int]]:
        """Contiguous element ranges of the flat buffer cut at the first parameter of each prefix in `boundaries` (given in
        storage order), returned LAST range first: the order in which a backward pass finishes them when the parameters are
        stored in forward order.  E.g. ("bottleneck.", "decoder.") -> [decoder.. end), [bottleneck.. decoder), [0, bottleneck)."""
        cuts = []
        for prefix in boundaries:
            offs = [self.offsets[k][0] for k in self.names if k.startswith(prefix)]
            if not offs:
                raise KeyError(f"no parameter starts with {prefix!r}")
            cuts.append(min(offs))
        if cuts != sorted(cuts) or (cuts and cuts[0] <= 0) or len(set(cuts)) != len(cuts):
            raise ValueError(f"bucket boundaries {list(boundaries)} are not in storage order")
        edges = [0] + cuts + [self.numel]
        return [(edges[i], edges[i + 1]) for i in range(len(edges) - 2, -1, -1)]

    def allreduce_bucket(self, bucket: Tuple[int, int], group=None):
        """Sum one range of the flat gradient over the ranks (a view: no copy)."""
        import torch.distributed as dist
        dist.all_reduce(self.grad[bucket[0]:bucket[1]], group=group)

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
                dtemb: Optional[torch.Tensor], stream: int, eps: float = 1e-5, sums: Optional[torch.Tensor] = None):
    N, D, H, W, C = x.shape
    if sums is None:
        sums = torch.empty(N, 2, dtype=torch.float64, device=x.hi.device)
    call("b2d_gn_silu_bwd", ptr(x.hi), ptr(x.lo), 1 if x.f16 else 0, ptr(dy.hi), ptr(dy.lo), 1 if dy.f16 else 0, ptr(dx.hi), ptr(dx.lo),
         1 if dx.f16 else 0, N, D * H * W, C, stats.data_ptr(), ptr(gamma), ptr(beta), float(eps), 1 if act else 0, sums.data_ptr(),
         dgamma.data_ptr(), dbeta.data_ptr(), ptr(dtemb), stream, launches=3)
    return sums


WGRAD_CONV3X3, WGRAD_LINEAR, WGRAD_CONVT2X2 = 0, 1, 2


def conv_wgrad(dy: Act, x: Act, dw: torch.Tensor, cout: int, cin: int, cin_off: int, stream: int, kind: int = WGRAD_CONV3X3,
               channels_last: bool = False):
    """dw (fp32, the reference's parameter layout) += the weight gradient of the input-channel block starting at cin_off.
    kind: WGRAD_CONV3X3 (dw [Cout, Cin_total, 3, 3]), WGRAD_LINEAR (dw [Cout, Cin_total]; Linear, Conv1d k1),
    WGRAD_CONVT2X2 (dw [Cin, Cout, 2, 2]; dy is the 2H x 2W gradient of the transposed conv's output)."""
    N, D, H, W, _ = x.shape
    up = 2 if kind == WGRAD_CONVT2X2 else 1
    assert D == 1 and dy.shape[:4] == (N, 1, H * up, W * up) and dw.dtype == torch.float32 and dw.is_contiguous()
    assert x.f16 == dy.f16 and (x.lo is None) == (dy.lo is None)
    # channels_last: dw is [rows, KH, KW, cols] (16-byte vector reductions) instead of the reference's [rows, cols, KH, KW]
    cin_total = dw.shape[0] if kind == WGRAD_CONVT2X2 else dw.shape[-1] if (channels_last and dw.dim() == 4) else dw.shape[1]
    call("b2d_conv_wgrad", kind, ptr(dy.hi), ptr(dy.lo), dy.C, ptr(x.hi), ptr(x.lo), x.C, N, H, W, cout, cin, cin_off, cin_total,
         dw.data_ptr(), 1 if channels_last else 0, 1 if x.f16 else 0, stream)


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


def attention_bwd(qkv: Act, out: Act, d_out: Act, d_qkv: Act, heads: int, stream: int, stats: Optional[torch.Tensor] = None):
    """Backward of engine's attention core (b2d_attention): d_qkv = (dq | dk | dv)."""
    N, D, H, W, C3 = qkv.shape
    C, T = C3 // 3, D * H * W
    assert out.shape == d_out.shape == (N, D, H, W, C) and d_qkv.shape == qkv.shape
    if stats is None:
        stats = torch.empty(N * heads * T * 2, dtype=torch.float32, device=qkv.hi.device)
    call("b2d_attention_bwd", ptr(qkv.hi), ptr(qkv.lo), ptr(out.hi), ptr(out.lo), ptr(d_out.hi), ptr(d_out.lo), ptr(d_qkv.hi), ptr(d_qkv.lo),
         stats.data_ptr(), N, T, C, heads, 1 if qkv.f16 else 0, stream, launches=3)


def pack_convT2x2_dgrad(w: torch.Tensor, device, split=False, f16=False) -> engine.PackedWeight:
    """Data gradient of nn.ConvTranspose2d k2 s2 (weight [Cin, Cout, 2, 2]): dX[p][ci] = sum_{ky,kx,co} dY[2p + (ky,kx)][co]
    W[ci][co][ky][kx] -- a stride-2 conv with a 2x2 window (ConvPlan(..., stride=2))."""
    ci, co = w.shape[:2]
    taps = [(0, ky, kx) for ky in range(2) for kx in range(2)]
    return engine.pack_weight(w.permute(0, 2, 3, 1).reshape(ci, 4, co), [co], taps, None, device, split, f16=f16)


class _no_tf32:
    """The parameter-space products of the folded attention projections are IEEE fp32 whatever the process-wide setting."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


# ------------------------------------------------------------------------------------------------ step-persistent state
class LiveOperand:
    """A packed conv operand (engine.PackedWeight) together with the recipe that rewrites it IN PLACE from its fp32 source
    parameter (b2d_pack_weight): after an optimizer step `refresh` brings the operand up to date while every plan that
    holds its address stays valid.  `parts`: one (src, elem_off, R1, R2, sr1, sr2, ntaps, st, cs, sc) per K segment."""

    def __init__(self, pw: engine.PackedWeight, parts):
        self.pw, self.parts = pw, parts

    def refresh(self, stream: int):
        pw = self.pw
        for i, (src, off, R1, R2, sr1, sr2, ntaps, st, cs, sc) in enumerate(self.parts):
            base = pw.w.data_ptr()
            lo = base + 2 * pw.kbase_lo[i] if pw.split else None
            call("b2d_pack_weight", src.data_ptr() + 4 * off, R1, R2, sr1, sr2, ntaps, st, cs, sc, base + 2 * pw.kbase[i], lo, pw.ktot,
                 pw.cin_pad[i], 1 if pw.f16 else 0, stream)


def live_conv2d(w: torch.Tensor, seg_sizes: Sequence[int], bias, device, split=True, channels_last=False) -> LiveOperand:
    """Forward operand of Conv2d 3x3 (w: a contiguous fp32 view of the parameter, [Cout, Cin, 3, 3], or [Cout, 3, 3, Cin] if
    channels_last -- the layout train.UNetTrainer keeps its conv parameters in)."""
    ref = w.permute(0, 3, 1, 2) if channels_last else w
    co, ci = ref.shape[:2]
    parts, c0 = [], 0
    for cs in seg_sizes:
        parts.append((w, c0, 1, co, 0, ci * 9, 9, ci, cs, 1) if channels_last else (w, c0 * 9, 1, co, 0, ci * 9, 9, 1, cs, 9))
        c0 += cs
    return LiveOperand(engine.pack_conv2d(ref, seg_sizes, bias, device, split=split), parts)


def live_conv2d_dgrad(w: torch.Tensor, seg: Tuple[int, int], device, split=True, channels_last=False) -> LiveOperand:
    """Data-gradient operand for input channels [seg0, seg1): rows = Cin of the segment, taps mirrored, columns = Cout."""
    ref = w.permute(0, 3, 1, 2) if channels_last else w
    co, ci = ref.shape[:2]
    c0, c1 = seg
    part = (w, c0 + 8 * ci, 1, c1 - c0, 0, 1, 9, -ci, co, ci * 9) if channels_last else (w, c0 * 9 + 8, 1, c1 - c0, 0, 9, 9, -1, co, ci * 9)
    return LiveOperand(pack_conv2d_dgrad(ref, seg, device, split=split), [part])


def live_linear(w: torch.Tensor, bias, device, transpose=False, split=True) -> LiveOperand:
    """nn.Linear / 1x1 weight [out, in]; transpose: the data-gradient operand [in, out]."""
    o, i = w.shape
    if transpose:
        return LiveOperand(engine.pack_linear(w.t().contiguous(), None, device, split=split), [(w, 0, 1, i, 0, 1, 1, 0, o, i)])
    return LiveOperand(engine.pack_linear(w, bias, device, split=split), [(w, 0, 1, o, 0, i, 1, 0, i, 1)])


def live_convT2x2(w: torch.Tensor, bias, device, split=True, channels_last=False) -> LiveOperand:
    """Forward operand of ConvTranspose2d k2 s2, weight [Cin, Cout, 2, 2] ([Cin, 2, 2, Cout] if channels_last): rows
    (phase, co), columns ci."""
    ref = w.permute(0, 3, 1, 2) if channels_last else w
    ci, co = ref.shape[:2]
    part = (w, 0, 4, co, co, 1, 1, 0, ci, co * 4) if channels_last else (w, 0, 4, co, 1, 4, 1, 0, ci, co * 4)
    return LiveOperand(engine.pack_convT2x2(ref, bias, device, split=split), [part])


def live_convT2x2_dgrad(w: torch.Tensor, device, split=True, channels_last=False) -> LiveOperand:
    ref = w.permute(0, 3, 1, 2) if channels_last else w
    ci, co = ref.shape[:2]
    part = (w, 0, 1, ci, 0, co * 4, 4, co, co, 1) if channels_last else (w, 0, 1, ci, 0, co * 4, 4, 1, co, 4)
    return LiveOperand(pack_convT2x2_dgrad(ref, device, split=split), [part])


class StepCache:
    """Buffers and plans of one training-step shape.  persistent=True: every activation, scratch tensor and conv plan is
    created on first use and reused by the following steps (accumulators are carved from one pool that `begin_step`
    clears with a single memset); persistent=False: allocate on every use (stand-alone block tests)."""

    POOL_DOUBLES = 1 << 16

    def __init__(self, device, persistent: bool, split: bool = True):
        self.dev, self.persistent, self.split = torch.device(device), persistent, split
        self.store: dict = {}
        self.pools: List[torch.Tensor] = []   # fp64 accumulator pools (a new one is added when the last is full)
        self.pool_used = 0

    def get(self, key, make):
        if not self.persistent:
            return make()
        if key not in self.store:
            self.store[key] = make()
        return self.store[key]

    def act(self, key, N, H, W, C, zero=False) -> Act:
        return self.get(key, lambda: new_act(N, 1, H, W, C, self.dev, split=self.split, zero=zero))

    def empty(self, key, shape, dtype=torch.float32) -> torch.Tensor:
        return self.get(key, lambda: torch.empty(shape, dtype=dtype, device=self.dev))

    def zeros64(self, key, shape) -> torch.Tensor:
        """An fp64 accumulator that is zero at the start of every step."""
        def make():
            if not self.persistent:
                return torch.zeros(shape, dtype=torch.float64, device=self.dev)
            n = 1
            for d in shape:
                n *= d
            if not self.pools or self.pool_used + n > self.pools[-1].numel():
                self.pools.append(torch.zeros(max(self.POOL_DOUBLES, n), dtype=torch.float64, device=self.dev))
                self.pool_used = 0
            t = self.pools[-1][self.pool_used:self.pool_used + n].view(shape)
            self.pool_used += n
            return t
        return self.get(key, make)

    def plan(self, key, make) -> ConvPlan:
        return self.get(key, make)

    def reset(self):
        """Forget every buffer and plan (a new batch shape); the accumulator pool is handed out again from a clean state."""
        self.store.clear()
        self.pools.clear()
        self.pool_used = 0

    def begin_step(self):
        for pool in self.pools:
            pool.zero_()


def _grad_dest(grads: Optional[dict], key: str, shape, dev) -> torch.Tensor:
    """Where a parameter gradient accumulates: the caller's (zeroed) destination, else a fresh zero tensor."""
    if grads is not None and key in grads:
        return grads[key]
    return torch.zeros(*shape, dtype=torch.float32, device=dev)


class DoubleBlockGrad:
    """Forward + backward of one DoubleBlock (unet/blocks.py:50-107) in the fp32-class mode, from the reference's own
    parameter tensors: conv{1,2}.weight [C, Cin, 3, 3] (no bias), norm{1,2}.{weight, bias}; `temb` (N, Cmid) is the per-sample
    time embedding added after block1 (blocks.py:100-103).  Inputs may be a channel concatenation (decoder blocks)."""

    def __init__(self, w1, g1, b1, w2, g2, b2, seg_sizes: Sequence[int], device="cuda", workspace: Optional[torch.Tensor] = None,
                 cache: Optional[StepCache] = None, name: str = "double", channels_last: bool = False):
        """channels_last: w1 / w2 (and the conv weight gradients `backward` produces) are [Cout, 3, 3, Cin]."""
        dev = torch.device(device)
        self.dev, self.name, self.cl = dev, name, channels_last
        self.seg_sizes = list(seg_sizes)
        f = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self.w1, self.w2 = f(w1), f(w2)
        self.cmid, self.cout = w1.shape[0], w2.shape[0]
        self.g1, self.b1, self.g2, self.b2 = f(g1), f(b1), f(g2), f(b2)
        self.cache = cache if cache is not None else StepCache(dev, False)
        sp = self.cache.split
        cl = channels_last
        self.l1 = live_conv2d(self.w1, self.seg_sizes, None, dev, sp, cl)
        self.l2 = live_conv2d(self.w2, [self.cmid], None, dev, sp, cl)
        self.ld2 = live_conv2d_dgrad(self.w2, (0, self.cmid), dev, sp, cl)
        self.ld1, c0 = [], 0
        for cs in self.seg_sizes:
            # the data gradient of a segment is only defined for engine-sized channel counts (the UNet's first layer has
            # 17 input channels and needs none: its input is data)
            self.ld1.append(live_conv2d_dgrad(self.w1, (c0, c0 + cs), dev, sp, cl) if cs % 64 == 0 else None)
            c0 += cs
        self.ws = workspace if workspace is not None else engine.new_workspace(dev)
        self.saved = None

    def operands(self) -> List[LiveOperand]:
        return [self.l1, self.l2, self.ld2] + [l for l in self.ld1 if l is not None]

    def forward(self, inputs: Sequence[Act], temb: Optional[torch.Tensor] = None, stats_out: Optional[torch.Tensor] = None) -> Act:
        """stats_out: (N, 2) fp64, receives (sum, sumsq) of the block's output (the next layer's GroupNorm statistics)."""
        N, D, H, W, _ = inputs[0].shape
        dev, s, c, k = self.dev, _lib.stream_ptr(), self.cache, self.name
        st1, st2 = c.zeros64(f"{k}.st1", (N, 2)), c.zeros64(f"{k}.st2", (N, 2))
        raw1, a1 = c.act(f"{k}.raw1", N, H, W, self.cmid), c.act(f"{k}.a1", N, H, W, self.cmid)
        raw2, out = c.act(f"{k}.raw2", N, H, W, self.cout), c.act(f"{k}.out", N, H, W, self.cout)
        c.plan(f"{k}.conv1", lambda: ConvPlan(list(inputs), self.l1.pw, raw1, cout=self.cmid, stats=st1, stats_cpg=self.cmid,
                                              workspace=self.ws)).run(s)
        row = c.get(f"{k}.row", lambda: torch.arange(N, dtype=torch.int32, device=dev)) if temb is not None else None
        tt = None
        if temb is not None:
            tt = c.empty(f"{k}.temb", (N, self.cmid))
            tt.copy_(temb)
        engine.gn_apply(raw1, a1, st1, self.cmid, self.g1, self.b1, True, s, temb=tt, temb_row=row, temb_row_stride=1)
        c.plan(f"{k}.conv2", lambda: ConvPlan([a1], self.l2.pw, raw2, cout=self.cout, stats=st2, stats_cpg=self.cout,
                                              workspace=self.ws)).run(s)
        engine.gn_apply(raw2, out, st2, self.cout, self.g2, self.b2, True, s, stats_out=stats_out)
        self.saved = (list(inputs), raw1, st1, a1, raw2, st2, temb is not None)
        return out

    def backward(self, d_out: Act, grads: Optional[dict] = None) -> dict:
        """Returns {'conv1.weight', 'norm1.weight', 'norm1.bias', 'conv2.weight', 'norm2.weight', 'norm2.bias', 'temb' (N, Cmid),
        'inputs': [Act or None per segment]} -- the gradients torch.autograd gives for the same block.  `grads`: zeroed
        destinations for the parameter gradients (same keys); missing ones are allocated."""
        inputs, raw1, st1, a1, raw2, st2, has_temb = self.saved
        N, D, H, W, _ = raw1.shape
        dev, s, c, k = self.dev, _lib.stream_ptr(), self.cache, self.name
        g = {"norm2.weight": _grad_dest(grads, "norm2.weight", (self.cout,), dev), "norm2.bias": _grad_dest(grads, "norm2.bias", (self.cout,), dev),
             "norm1.weight": _grad_dest(grads, "norm1.weight", (self.cmid,), dev), "norm1.bias": _grad_dest(grads, "norm1.bias", (self.cmid,), dev),
             "conv2.weight": _grad_dest(grads, "conv2.weight", (self.cout, 3, 3, self.cmid) if self.cl else (self.cout, self.cmid, 3, 3), dev),
             "conv1.weight": _grad_dest(grads, "conv1.weight", (self.cmid, 3, 3, sum(self.seg_sizes)) if self.cl
                                        else (self.cmid, sum(self.seg_sizes), 3, 3), dev), "temb": None}
        if has_temb:
            g["temb"] = c.empty(f"{k}.dtemb", (N, self.cmid))
            g["temb"].zero_()
        d_raw2 = c.act(f"{k}.d_raw2", N, H, W, self.cout)
        gn_silu_bwd(raw2, d_out, d_raw2, st2, self.g2, self.b2, True, g["norm2.weight"], g["norm2.bias"], None, s,
                    sums=c.empty(f"{k}.sums2", (N, 2), torch.float64))
        conv_wgrad(d_raw2, a1, g["conv2.weight"], self.cout, self.cmid, 0, s, channels_last=self.cl)
        d_a1 = c.act(f"{k}.d_a1", N, H, W, self.cmid)
        c.plan(f"{k}.dgrad2", lambda: ConvPlan([d_raw2], self.ld2.pw, d_a1, cout=self.cmid, workspace=self.ws)).run(s)
        d_raw1 = c.act(f"{k}.d_raw1", N, H, W, self.cmid)
        gn_silu_bwd(raw1, d_a1, d_raw1, st1, self.g1, self.b1, True, g["norm1.weight"], g["norm1.bias"], g["temb"], s,
                    sums=c.empty(f"{k}.sums1", (N, 2), torch.float64))
        d_inputs, c0 = [], 0
        for i, (x, cs, ld) in enumerate(zip(inputs, self.seg_sizes, self.ld1)):
            conv_wgrad(d_raw1, x, g["conv1.weight"], self.cmid, cs, c0, s, channels_last=self.cl)
            if ld is None:
                d_inputs.append(None)
            else:
                dx = c.act(f"{k}.dx{i}", N, H, W, x.C)
                c.plan(f"{k}.dgrad1.{i}", lambda: ConvPlan([d_raw1], ld.pw, dx, cout=cs, workspace=self.ws)).run(s)
                d_inputs.append(dx)
            c0 += cs
        g["inputs"] = d_inputs
        return g


class AttentionGrad:
    """Forward + backward of SelfAttention2d (unet/blocks.py:170-235): y = x + proj_out(out_proj(softmax(q k^T / sqrt(d)) v)),
    (q | k | v) = in_proj(GroupNorm(1, C)(x)).  The two C x C output projections run folded (W = Wp Wo, b = Wp bo + bp) as in
    the sampling path; their separate gradients follow from the folded ones by C x C parameter-space products."""

    def __init__(self, c: int, heads: int, p: Dict[str, torch.Tensor], device="cuda", workspace: Optional[torch.Tensor] = None,
                 cache: Optional[StepCache] = None, name: str = "attn"):
        dev = torch.device(device)
        f = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self.dev, self.c, self.heads, self.name = dev, c, heads, name
        self.g, self.b = f(p["norm.weight"]), f(p["norm.bias"])
        self.w_in, self.b_in = f(p["mha.in_proj_weight"]), f(p["mha.in_proj_bias"])
        self.wo, self.bo = f(p["mha.out_proj.weight"]), f(p["mha.out_proj.bias"])
        self.wp3, self.bp = f(p["proj_out.weight"]), f(p["proj_out.bias"])
        self.wp = self.wp3[:, :, 0]
        self.w_out = torch.empty(c, c, dtype=torch.float32, device=dev)   # Wp Wo, recomputed by refresh_folded()
        self.b_out = torch.empty(c, dtype=torch.float32, device=dev)
        self.refresh_folded()
        self.cache = cache if cache is not None else StepCache(dev, False)
        sp = self.cache.split
        self.l_in = live_linear(self.w_in, self.b_in, dev, split=sp)
        self.l_out = live_linear(self.w_out, self.b_out, dev, split=sp)
        self.l_in_t = live_linear(self.w_in, None, dev, transpose=True, split=sp)
        self.l_out_t = live_linear(self.w_out, None, dev, transpose=True, split=sp)
        self.ws = workspace if workspace is not None else engine.new_workspace(dev)
        self.saved = None

    def refresh_folded(self):
        # fp32 products (IEEE, no TF32): the folded operand is rounded to bf16 hi + lo (2^-16) right after
        with _no_tf32():
            torch.matmul(self.wp, self.wo, out=self.w_out)
            torch.addmv(self.bp, self.wp, self.bo, out=self.b_out)

    def operands(self) -> List[LiveOperand]:
        return [self.l_in, self.l_out, self.l_in_t, self.l_out_t]

    def forward(self, x: Act, st_x: torch.Tensor) -> Act:
        """st_x: (N, 2) fp64 (sum, sumsq) of x per sample."""
        N, D, H, W, c = x.shape
        dev, s, ch, k = self.dev, _lib.stream_ptr(), self.cache, self.name
        xn = ch.act(f"{k}.xn", N, H, W, c)
        engine.gn_apply(x, xn, st_x, c, self.g, self.b, False, s)
        qkv = ch.act(f"{k}.qkv", N, H, W, 3 * c)
        ch.plan(f"{k}.in", lambda: ConvPlan([xn], self.l_in.pw, qkv, cout=3 * c, workspace=self.ws)).run(s)
        ao = ch.act(f"{k}.ao", N, H, W, c)
        call("b2d_attention", ptr(qkv.hi), ptr(qkv.lo), ptr(ao.hi), ptr(ao.lo), N, H * W, c, self.heads, 0, s)
        y = ch.act(f"{k}.y", N, H, W, c)
        ch.plan(f"{k}.out", lambda: ConvPlan([ao], self.l_out.pw, y, cout=c, residual=x, workspace=self.ws)).run(s)
        self.saved = (x, st_x, xn, qkv, ao)
        return y

    def backward(self, d_y: Act, grads: Optional[dict] = None) -> Tuple[dict, Act]:
        x, st_x, xn, qkv, ao = self.saved
        N, D, H, W, c = x.shape
        dev, s, ch, k = self.dev, _lib.stream_ptr(), self.cache, self.name
        d_wout, d_bout = ch.empty(f"{k}.d_wout", (c, c)), ch.empty(f"{k}.d_bout", (c,))
        d_wout.zero_()
        d_bout.zero_()
        conv_wgrad(d_y, ao, d_wout, c, c, 0, s, kind=WGRAD_LINEAR)
        channel_sum(d_y, d_bout, c, s)
        d_ao = ch.act(f"{k}.d_ao", N, H, W, c)
        ch.plan(f"{k}.out_t", lambda: ConvPlan([d_y], self.l_out_t.pw, d_ao, cout=c, workspace=self.ws)).run(s)
        d_qkv = ch.act(f"{k}.d_qkv", N, H, W, 3 * c)
        attention_bwd(qkv, ao, d_ao, d_qkv, self.heads, s, stats=ch.empty(f"{k}.attn_stats", (N * self.heads * H * W * 2,)))
        g = {"mha.in_proj_weight": _grad_dest(grads, "mha.in_proj_weight", (3 * c, c), dev),
             "mha.in_proj_bias": _grad_dest(grads, "mha.in_proj_bias", (3 * c,), dev),
             "norm.weight": _grad_dest(grads, "norm.weight", (c,), dev), "norm.bias": _grad_dest(grads, "norm.bias", (c,), dev)}
        conv_wgrad(d_qkv, xn, g["mha.in_proj_weight"], 3 * c, c, 0, s, kind=WGRAD_LINEAR)
        channel_sum(d_qkv, g["mha.in_proj_bias"], 3 * c, s)
        d_xn = ch.act(f"{k}.d_xn", N, H, W, c)
        ch.plan(f"{k}.in_t", lambda: ConvPlan([d_qkv], self.l_in_t.pw, d_xn, cout=c, workspace=self.ws)).run(s)
        d_x = ch.act(f"{k}.d_x", N, H, W, c)
        gn_silu_bwd(x, d_xn, d_x, st_x, self.g, self.b, False, g["norm.weight"], g["norm.bias"], None, s,
                    sums=ch.empty(f"{k}.sums", (N, 2), torch.float64))
        add_acts(d_x, d_y, d_x, s)
        # W = Wp Wo, b = Wp bo + bp  ->  dWp = dW Wo^T + db bo^T, dWo = Wp^T dW, dbo = Wp^T db, dbp = db
        with _no_tf32():
            out = {"proj_out.weight": torch.addmm(torch.outer(d_bout, self.bo), d_wout, self.wo.t())[:, :, None],
                   "proj_out.bias": d_bout, "mha.out_proj.weight": self.wp.t() @ d_wout, "mha.out_proj.bias": self.wp.t() @ d_bout}
        for key, v in out.items():
            if grads is not None and key in grads:
                grads[key].copy_(v)
                g[key] = grads[key]
            else:
                g[key] = v
        return g, d_x


class UNetTrainer:
    """One optimisation step of the conditioned eps-prediction UNet (helper.py:420-431, predictor.py:722-748):

        x_t = q_sample(x_start, t, noise);  pred = UNet(cat[x_t, cond, feats], t);  loss = criterion(pred, noise)
        loss.backward();  optimizer.step()                                     (torch.optim.Adam, train.py:144-148)

    in the fp32-class mode (bf16 hi + lo operands, three MMA passes): every layer's forward saves what its backward needs,
    the backward walks the UNet (unet/models.py:131-188) in reverse -- final_conv, decoder levels (attention, DoubleBlock over
    cat[skip, up], ConvTranspose2d + GroupNorm + SiLU), bottleneck, encoder levels (max-pool + GroupNorm + SiLU, attention,
    DoubleBlock) -- accumulating each parameter's gradient directly in FlatAdam's flat buffer, which one launch then applies.
    Activations, scratch and conv plans are created by the first step of a shape and reused (StepCache); after the update
    every packed operand is rewritten in place from the new parameters (LiveOperand).
    The sinusoid -> time_mlp -> per-block Linear chain ((N, 64) -> (N, 256) -> (N, Cmid): a few kFLOP) is evaluated and
    differentiated with torch ops on the GPU; everything that touches a feature map is libb2d."""

    _ATTN = ("norm.weight", "norm.bias", "mha.in_proj_weight", "mha.in_proj_bias", "mha.out_proj.weight", "mha.out_proj.bias",
             "proj_out.weight", "proj_out.bias")
    _DOUBLE = {"conv1.weight": "block1.conv.weight", "conv2.weight": "block2.conv.weight", "norm1.weight": "block1.norm.weight",
               "norm1.bias": "block1.norm.bias", "norm2.weight": "block2.norm.weight", "norm2.bias": "block2.norm.bias"}

    def __init__(self, state_dict: Dict[str, torch.Tensor], *, in_channels=17, out_channels=8, features=(64, 128, 256, 512, 1024),
                 attention: str = "", time_embedding_dim: Optional[int] = 64, num_timesteps: int = 1000, lr: float = 1e-4,
                 weight_decay: float = 0.0, device="cuda", precision: str = "fp32x", **_ignored):
        """precision: "fp32x" (bf16 hi + lo operands and activations, three MMA passes: the parity mode) or "bf16" (single
        bf16 operands / activations / gradients, fp32 accumulate: a third of the MMA work, gradients to ~1e-2)."""
        from .scheduler import B200Scheduler
        from .synth import attention_heads
        if not torch.cuda.is_available():
            raise RuntimeError("UNetTrainer runs on a CUDA device only (no CPU fallback)")
        # the configurations B200UNet supports (the shipped model): anything else is refused, not approximated
        bad = {k: v for k, v in _ignored.items() if (k, v) not in (("kernel_size", 3), ("padding_mode", "zeros"), ("activation", "silu"),
                                                                    ("final_activation", None), ("dropout", 0.0))}
        if bad:
            raise NotImplementedError(f"UNetTrainer: unsupported model arguments {bad}")
        self.dev = torch.device(device)
        self.in_channels, self.out_channels, self.features = in_channels, out_channels, list(features)
        self.time_dim = time_embedding_dim
        self.heads = attention_heads(attention, len(self.features))
        # 4-D parameters (Conv2d [Cout, Cin, 3, 3], ConvTranspose2d [Cin, Cout, 2, 2]) are kept with their second index last
        # ([Cout, 3, 3, Cin], [Cin, 2, 2, Cout]): the order of the packed operands' K axis and of the weight-gradient
        # kernel's accumulator rows, so refresh and wgrad move contiguous runs; P / G / state_dict present the
        # reference's layout as views
        self.opt = FlatAdam({k: (v.permute(0, 2, 3, 1).contiguous() if v.dim() == 4 else v) for k, v in state_dict.items()},
                            lr=lr, weight_decay=weight_decay, device=self.dev)
        self.scheduler = B200Scheduler(num_timesteps=num_timesteps, device=self.dev)
        self.ws = engine.new_workspace(self.dev)
        if precision not in ("fp32x", "bf16"):
            raise ValueError(f"UNetTrainer precision must be 'fp32x' or 'bf16', got {precision!r}")
        self.precision, self.split = precision, precision == "fp32x"
        self.cache = StepCache(self.dev, True, self.split)
        self._shape = None
        self._layers = None
        self._graph = None
        self._comm_stream = None
        self._fb_launches = 0
        self._build_layers()

    # -------------------------------------------------------------------------------- parameters
    def _Pk(self, name: str) -> torch.Tensor:
        return self.opt.view(self.opt.param, name)

    def _Gk(self, name: str) -> torch.Tensor:
        return self.opt.view(self.opt.grad, name)

    @staticmethod
    def _ref_layout(t: torch.Tensor) -> torch.Tensor:
        return t.permute(0, 3, 1, 2) if t.dim() == 4 else t

    def P(self, name: str) -> torch.Tensor:
        """The parameter in the reference's layout (a view of the flat buffer)."""
        return self._ref_layout(self._Pk(name))

    def G(self, name: str) -> torch.Tensor:
        """Its gradient, same layout."""
        return self._ref_layout(self._Gk(name))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self._ref_layout(v) for k, v in self.opt.state_dict().items()}

    def _build_layers(self):
        """Operand forms of the parameters (built once; `refresh_operands` rewrites them in place after each update)."""
        dev, ws, f, ch = self.dev, self.ws, self.features, self.cache
        L = {}

        def double(p, segs):
            L[p] = DoubleBlockGrad(self._Pk(f"{p}.block1.conv.weight"), self.P(f"{p}.block1.norm.weight"), self.P(f"{p}.block1.norm.bias"),
                                   self._Pk(f"{p}.block2.conv.weight"), self.P(f"{p}.block2.norm.weight"), self.P(f"{p}.block2.norm.bias"),
                                   segs, dev, workspace=ws, cache=ch, name=p, channels_last=True)

        def attn(p, c, heads):
            L[p] = AttentionGrad(c, heads, {n: self.P(f"{p}.{n}") for n in self._ATTN}, dev, ws, cache=ch, name=p)

        cin = self.in_channels
        for lvl, c in enumerate(f):
            double(f"encoder.{lvl}.0", [cin])
            if self.heads[lvl] is not None:
                attn(f"encoder.{lvl}.1", c, self.heads[lvl])
            cin = c
        double("bottleneck", [f[-1]])
        rheads = list(reversed(self.heads))
        for lvl, c in enumerate(reversed(f)):
            w = self._Pk(f"decoder.{lvl}.0.conv.weight")
            L[f"decoder.{lvl}.0"] = (live_convT2x2(w, self.P(f"decoder.{lvl}.0.conv.bias"), dev, self.split, True),
                                     live_convT2x2_dgrad(w, dev, self.split, True))
            double(f"decoder.{lvl}.1", [c, c])
            if rheads[lvl] is not None:
                attn(f"decoder.{lvl}.2", c, rheads[lvl])
        wf = self._Pk("final_conv.weight")
        L["final_conv"] = (live_conv2d(wf, [f[0]], self.P("final_conv.bias"), dev, self.split, True),
                           live_conv2d_dgrad(wf, (0, f[0]), dev, self.split, True))
        self._layers = L

    def refresh_operands(self):
        """Bring every packed operand up to date with the flat parameter buffer (after FlatAdam.step)."""
        s = _lib.stream_ptr()
        for layer in self._layers.values():
            if isinstance(layer, AttentionGrad):
                layer.refresh_folded()
            ops = layer.operands() if hasattr(layer, "operands") else list(layer)
            for op in ops:
                op.refresh(s)

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
        for _ in self._fb_phases(x, t, target):
            pass
        return self._fb_result

    def _fb_phases(self, x: torch.Tensor, t: torch.Tensor, target: torch.Tensor):
        """forward_backward as a generator that pauses where a contiguous block of the flat gradient is final -- after the
        decoder's backward ("decoder": decoder.* and final_conv.*, the tail of the buffer) and after the bottleneck's
        ("bottleneck") -- so that training_step can start reducing those blocks while the rest of the backward runs.
        The stream is read again after every pause (each phase may be captured on its own)."""
        L, dev, f, s = self._layers, self.dev, self.features, _lib.stream_ptr()
        N, _, h, w = x.shape
        nl = len(f)
        if h % (1 << nl) or w % (1 << nl):
            raise ValueError(f"UNetTrainer: input {h}x{w} must be divisible by {1 << nl}")
        if self._shape != (N, h, w):   # a new shape: new buffers and plans (the layers keep their operands)
            self.cache.reset()
            self._shape = (N, h, w)
            if self._graph is not None:
                self._graph["fb"] = None   # captured over the old buffers
        ch = self.cache
        ch.begin_step()
        self.opt.zero_grad()
        G = self._Gk   # kernel layout: the weight-gradient kernel accumulates channels-last
        temb, leaves = self._time_forward(t) if self.time_dim is not None else ({}, {})
        x_in = ch.act("x_in", N, h, w, engine.pad64(self.in_channels), zero=True)
        engine.planar_to_cl(x.contiguous().float(), x_in, N, self.in_channels, h * w, 0, None, s)
        # ---- forward (models.py:144-186)
        a, H, Wd = x_in, h, w
        skips, pools = [], []
        for lvl, c in enumerate(f):
            p = f"encoder.{lvl}.0"
            st = ch.zeros64(f"{p}.st_out", (N, 2)) if self.heads[lvl] is not None else None
            a = L[p].forward([a], temb[p].detach() if p in temb else None, stats_out=st)
            if st is not None:
                a = L[f"encoder.{lvl}.1"].forward(a, st)
            skips.append(a)
            q = f"encoder.{lvl}.2"
            praw, pst = ch.act(f"{q}.raw", N, H // 2, Wd // 2, c), ch.zeros64(f"{q}.st", (N, 2))
            engine.maxpool_stats(a, praw, pst, s)
            pa = ch.act(f"{q}.out", N, H // 2, Wd // 2, c)
            engine.gn_apply(praw, pa, pst, c, self.P(f"{q}.norm.weight"), self.P(f"{q}.norm.bias"), True, s)
            pools.append((a, praw, pst))
            a, H, Wd = pa, H // 2, Wd // 2
        a = L["bottleneck"].forward([a], temb["bottleneck"].detach() if temb else None)
        ups = []
        rheads = list(reversed(self.heads))
        for lvl, c in enumerate(reversed(f)):
            q = f"decoder.{lvl}.0"
            raw, ust = ch.act(f"{q}.raw", N, 2 * H, 2 * Wd, c), ch.zeros64(f"{q}.st", (N, 2))
            ch.plan(f"{q}.conv", lambda: ConvPlan([a], L[q][0].pw, raw, cout=c, nphase=4, stats=ust, stats_cpg=c, workspace=self.ws)).run(s)
            up = ch.act(f"{q}.out", N, 2 * H, 2 * Wd, c)
            engine.gn_apply(raw, up, ust, c, self.P(f"{q}.norm.weight"), self.P(f"{q}.norm.bias"), True, s)
            ups.append((a, raw, ust))
            H, Wd = 2 * H, 2 * Wd
            p = f"decoder.{lvl}.1"
            st = ch.zeros64(f"{p}.st_out", (N, 2)) if rheads[lvl] is not None else None
            a = L[p].forward([skips[nl - 1 - lvl], up], temb[p].detach() if p in temb else None, stats_out=st)
            if st is not None:
                a = L[f"decoder.{lvl}.2"].forward(a, st)
        pred = ch.empty("pred", (N, self.out_channels, h, w))
        final_in = a
        ch.plan("final_conv", lambda: ConvPlan([final_in], L["final_conv"][0].pw, pred, cout=self.out_channels, out_mode=1,
                                               out_cstride=self.out_channels, workspace=self.ws)).run(s)
        # ---- criterion (metrics.py:337-402)
        loss, _, d_pred = nmse_loss(pred, target.to(dev))
        # ---- backward
        d_temb = {}
        oc, c0 = self.out_channels, f[0]
        d = ch.act("d_pred", N, h, w, engine.pad64(oc), zero=True)
        engine.planar_to_cl(d_pred, d, N, oc, h * w, 0, None, s)
        channel_sum(d, G("final_conv.bias"), oc, s)
        conv_wgrad(d, final_in, G("final_conv.weight"), oc, c0, 0, s, channels_last=True)
        da = ch.act("d_final_in", N, h, w, c0)
        ch.plan("final_conv.dgrad", lambda: ConvPlan([d], L["final_conv"][1].pw, da, cout=c0, workspace=self.ws)).run(s)
        d_skips = [None] * nl

        def double_bwd(p, d_out):
            g = L[p].backward(d_out, {k: G(f"{p}.{v}") for k, v in self._DOUBLE.items()})
            if g["temb"] is not None:
                d_temb[p] = g["temb"]
            return g["inputs"]

        def attention_bwd_(p, d_out):
            return L[p].backward(d_out, {n: G(f"{p}.{n}") for n in self._ATTN})[1]

        for lvl in range(nl - 1, -1, -1):
            c = f[nl - 1 - lvl]
            if rheads[lvl] is not None:
                da = attention_bwd_(f"decoder.{lvl}.2", da)
            d_skips[nl - 1 - lvl], d_up = double_bwd(f"decoder.{lvl}.1", da)
            q = f"decoder.{lvl}.0"
            x_lo, raw, ust = ups[lvl]
            _, _, H2, W2, _ = raw.shape
            d_raw = ch.act(f"{q}.d_raw", N, H2, W2, c)
            gn_silu_bwd(raw, d_up, d_raw, ust, self.P(f"{q}.norm.weight"), self.P(f"{q}.norm.bias"), True, G(f"{q}.norm.weight"),
                        G(f"{q}.norm.bias"), None, s, sums=ch.empty(f"{q}.sums", (N, 2), torch.float64))
            conv_wgrad(d_raw, x_lo, G(f"{q}.conv.weight"), c, 2 * c, 0, s, kind=WGRAD_CONVT2X2, channels_last=True)
            channel_sum(d_raw, G(f"{q}.conv.bias"), c, s)
            da = ch.act(f"{q}.d_in", N, H2 // 2, W2 // 2, 2 * c)
            d_in = da
            ch.plan(f"{q}.dgrad", lambda: ConvPlan([d_raw], L[q][1].pw, d_in, cout=2 * c, stride=2, workspace=self.ws)).run(s)
        def time_bwd(prefixes, last):
            """Backward of the time-embedding chain for these blocks' Linear layers (their gradients live inside the blocks'
            region of the flat buffer); the shared time_mlp accumulates in the leaves until the last call."""
            if not temb:
                return
            ps = [p for p in prefixes if p in d_temb]
            torch.autograd.backward([temb[p] for p in ps], [d_temb[p] for p in ps], retain_graph=not last)
            for p in ps:
                for k in (f"{p}.time_mlp.1.weight", f"{p}.time_mlp.1.bias"):
                    G(k).copy_(leaves[k].grad)
            if last:
                for k in ("time_mlp.0.weight", "time_mlp.0.bias", "time_mlp.2.weight", "time_mlp.2.bias"):
                    G(k).copy_(leaves[k].grad)

        time_bwd([f"decoder.{l}.1" for l in range(nl)], False)
        yield "decoder"
        s = _lib.stream_ptr()
        da = double_bwd("bottleneck", da)[0]
        time_bwd(["bottleneck"], False)
        yield "bottleneck"
        s = _lib.stream_ptr()
        for lvl in range(nl - 1, -1, -1):
            c = f[lvl]
            q = f"encoder.{lvl}.2"
            skip, praw, pst = pools[lvl]
            d_praw = ch.act(f"{q}.d_raw", *praw.shape[:1], *praw.shape[2:])
            gn_silu_bwd(praw, da, d_praw, pst, self.P(f"{q}.norm.weight"), self.P(f"{q}.norm.bias"), True, G(f"{q}.norm.weight"),
                        G(f"{q}.norm.bias"), None, s, sums=ch.empty(f"{q}.sums", (N, 2), torch.float64))
            da = ch.act(f"{q}.d_in", *skip.shape[:1], *skip.shape[2:])
            maxpool_bwd(skip, d_praw, da, s)
            add_acts(da, d_skips[lvl], da, s)
            if self.heads[lvl] is not None:
                da = attention_bwd_(f"encoder.{lvl}.1", da)
            da = double_bwd(f"encoder.{lvl}.0", da)[0]
        time_bwd([f"encoder.{l}.0" for l in range(nl)], True)
        self._fb_result = (loss, pred)

    def _q_sample_phases(self, x_start, cond, feats, t, noise):
        x_t = self.scheduler.q_sample(x_start, t, noise)                                     # diffusion.py:78-101
        yield from self._fb_phases(torch.cat([x_t, cond, feats], dim=1), t, noise)           # predictor.py:731-741

    def _buckets(self):
        """The flat gradient in the order its blocks become final during the backward: [decoder.* + final_conv.*],
        [bottleneck.*], [time_mlp.* + encoder.*] (the parameters are stored in the reference's order: time_mlp, encoder,
        bottleneck, decoder, final_conv)."""
        return self.opt.backward_order_buckets(("bottleneck.", "decoder."))

    def training_step(self, x_start: torch.Tensor, cond: torch.Tensor, feats: torch.Tensor, t: torch.Tensor, noise: torch.Tensor,
                      group=None, use_graph: bool = True):
        """q_sample -> forward -> loss -> backward -> gradient all-reduce -> Adam -> operand refresh.  Returns (loss, pred).
        The backward pauses where a block of the flat gradient is final (decoder + final_conv, then the bottleneck); with more
        than one rank those blocks are all-reduced on a side stream while the rest of the backward runs, and only the last
        block (time MLP + encoder, a fifth of the parameters) is reduced after it.
        use_graph: from the second step of a batch shape the ~440 launches of q_sample .. backward replay as three CUDA
        graphs (one per phase) and the ~90 of the operand refresh as a fourth, over the StepCache's static buffers; the
        all-reduces and the Adam launch (whose bias corrections change every step) stay eager.  The returned tensors are
        then overwritten by the next step."""
        import torch.distributed as dist
        dev = self.dev
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        ins = [x_start, cond, feats, t, noise]
        main = torch.cuda.current_stream(dev)
        if world > 1 and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=dev)
        buckets = self._buckets()

        def reduce_async(i):
            """all-reduce bucket i on the side stream once everything issued so far on the main stream is done"""
            if world == 1:
                return
            self._comm_stream.wait_stream(main)
            with torch.cuda.stream(self._comm_stream):
                self.opt.allreduce_bucket(buckets[i], group)

        gr = None
        if use_graph:
            key = tuple(tuple(v.shape) for v in ins)
            if self._graph is None or self._graph["key"] != key:
                self._graph = {"key": key, "ins": [torch.empty(v.shape, dtype=v.dtype, device=dev) for v in ins], "fb": None}
            gr = self._graph
            for dst, v in zip(gr["ins"], ins):
                dst.copy_(v, non_blocking=True)
            if gr["fb"] is None and self._shape == (x_start.shape[0], x_start.shape[2], x_start.shape[3]):
                # buffers and plans of this shape exist (an eager step ran): capture the three phases and the refresh
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    gen = self._q_sample_phases(*gr["ins"])
                    graphs, pool = [], None
                    for _ in range(3):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=side, pool=pool):
                            next(gen, None)
                        pool = g.pool()
                        graphs.append(g)
                    assert next(gen, "done") == "done"
                    gr["out"] = self._fb_result
                    rf = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(rf, stream=side, pool=pool):
                        self.refresh_operands()
                main.wait_stream(side)
                gr["fb"], gr["refresh"] = graphs, rf
        if gr is not None and gr["fb"] is not None:
            for i, g in enumerate(gr["fb"]):
                g.replay()
                if i < 2:
                    reduce_async(i)
            _lib.launch_count += self._fb_launches
            loss, pred = gr["out"]
        else:
            n0 = _lib.launch_count
            gen = self._q_sample_phases(*(gr["ins"] if gr is not None else [v.to(dev) for v in ins]))
            for i, _ in enumerate(gen):
                reduce_async(i)
            loss, pred = self._fb_result
            self._fb_launches = _lib.launch_count - n0
        scale = 1.0
        if world > 1:
            self.opt.allreduce_bucket(buckets[2], group)            # the last block, on the main stream
            main.wait_stream(self._comm_stream)
            scale = 1.0 / world
        self.opt.step(grad_scale=scale)
        if gr is not None and gr["fb"] is not None:
            gr["refresh"].replay()
        else:
            self.refresh_operands()
        return loss, pred


class LatentDiffusionTrainer:
    """The 'latent-diffusion' branch of the reference's training loop body (helper.py:277-430, default losses: the noise
    criterion only -- physics / velocity losses are out of scope), from the fields a data loader yields:

        target_latents = predictor.encode_target(targets, velocity_2d)             # frozen E3D      (helper.py:288)
        noise = randn_like(target_latents);  t = randint(0, T, (N,))                #                 (:299, predictor.py:736)
        preds, noise = predictor(img, velocity_2d, x_start=target_latents, noise)   # frozen E2D conditioning + EDT
                                                                                    # features + q_sample + UNet (:636-751)
        loss = criterion(preds, noise);  zero_grad();  loss.backward();  optimizer.step()            (helper.py:428-430)

    `predictor` supplies the frozen VAE passes (its E2D / E3D programs, EDT and bilinear kernels) and is not updated;
    `unet` is the `UNetTrainer` that owns the parameters being optimised (built from the predictor's UNet state unless
    given).  `sync_predictor()` hands the trained parameters back to the predictor's sampling UNet."""

    def __init__(self, predictor, unet: Optional[UNetTrainer] = None, *, unet_state: Optional[Dict[str, torch.Tensor]] = None,
                 lr: float = 1e-4, weight_decay: float = 0.0, precision: str = "fp32x"):
        self.predictor = predictor
        if unet is None:
            if unet_state is None:
                raise ValueError("LatentDiffusionTrainer needs the UNet parameters: pass `unet` (a UNetTrainer) or `unet_state`")
            m = predictor.model
            unet = UNetTrainer(unet_state, in_channels=m.in_channels, out_channels=m.out_channels, features=tuple(m.features),
                               attention=m.attention, time_embedding_dim=m.time_embedding_dim, num_timesteps=predictor.num_timesteps,
                               lr=lr, weight_decay=weight_decay, device=predictor.device, precision=precision)
        self.unet = unet

    def latents(self, img: torch.Tensor, velocity_2d: torch.Tensor, targets: torch.Tensor):
        """(x_start, cond, feats), each fp32 (N, C, h, w): the frozen-VAE side of the step.  They depend on the sample only,
        not on the UNet: a data pipeline may compute them once per sample and cache them."""
        p = self.predictor
        x_start = p.encode_target(targets.to(p.device), velocity_2d)
        cond, feats = p.conditioning_latents(img, velocity_2d)
        return x_start.reshape(cond.shape).contiguous(), cond, feats

    def train_step(self, img: torch.Tensor, velocity_2d: torch.Tensor, targets: torch.Tensor, *, t: Optional[torch.Tensor] = None,
                   noise: Optional[torch.Tensor] = None, group=None, use_graph: bool = True):
        """One iteration of the loop body.  t / noise inject the reference's randint / randn_like draws (parity tests).
        Returns (loss, noise_pred, noise)."""
        x_start, cond, feats = self.latents(img, velocity_2d, targets)
        dev = x_start.device
        noise = torch.randn_like(x_start) if noise is None else noise.to(dev, torch.float32).reshape(x_start.shape)
        if t is None:
            t = torch.randint(0, self.predictor.num_timesteps, (x_start.shape[0],), device=dev).long()
        loss, pred = self.unet.training_step(x_start, cond, feats, t.to(dev).long(), noise, group=group, use_graph=use_graph)
        return loss, pred, noise

    def sync_predictor(self):
        """Load the optimised parameters into the predictor's sampling UNet (repacks its operands, rebuilds the time table)."""
        self.predictor.model.load_state_dict({k: v.detach().clone() for k, v in self.unet.state_dict().items()})
        self.predictor._session = None  # programs hold the old operands' addresses
