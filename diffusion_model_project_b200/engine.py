"""Host-side plumbing over the C ABI: weight repacking (reference layouts -> K-major bf16 GEMM
operands), conv plans, channels-last activation buffers and launch programs.

PyTorch is used for device memory, streams and one-off weight preprocessing only; every
per-step operation is a libb2d kernel launched through diffusion_model_project_b200._lib.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, call, ptr

BF16 = torch.bfloat16
# precision modes of the modules: 16-bit operands in IEEE fp16 ("f16", the default: 10-bit mantissa) or bf16 ("bf16"), or the
# fp32-class mode "fp32x" (bf16 hi + lo operands, three tensor-core passes)
PRECISIONS = ("f16", "bf16", "fp32x")


def pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def split_hi_lo(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> (bf16 hi, bf16 lo) with hi + lo ~= w to ~2^-16 relative (the "fp32x" mode)."""
    hi = w.to(BF16)
    lo = (w - hi.float()).to(BF16)
    return hi, lo


@dataclass
class Act:
    """Channels-last activation [N, D, H, W, C], 16-bit storage (C multiple of 64; `lo` only in fp32x mode).

    f16=True: the storage holds IEEE fp16 rather than bf16 (the tensors are allocated as torch.bfloat16 either way: torch
    only provides the memory).  In the "f16" precision mode every activation is fp16; in the "bf16" mode only the raw
    pre-GroupNorm conv outputs / residual streams are (read by gn_apply and residual adds, never by an MMA); the fp32x
    mode (bf16 hi + lo) has none."""
    hi: torch.Tensor
    lo: Optional[torch.Tensor] = None
    f16: bool = False

    def as_fmt(self, f16: bool) -> "Act":
        """The same storage after an in-place GroupNorm apply (which rewrites it in the operand format)."""
        return Act(self.hi, self.lo, bool(f16) and self.lo is None)

    @property
    def shape(self):
        return tuple(self.hi.shape)

    @property
    def C(self):
        return self.hi.shape[-1]


def new_act(N, D, H, W, C, device, split=False, zero=False, f16=False) -> Act:
    mk = torch.zeros if zero else torch.empty
    hi = mk((N, D, H, W, C), dtype=BF16, device=device)
    lo = mk((N, D, H, W, C), dtype=BF16, device=device) if split else None
    return Act(hi, lo, bool(f16) and not split)


# ------------------------------------------------------------------------------------------------
# weight packing: [rows][K] with K = sum_seg ntaps * cin_pad[seg]; k = kbase[seg] + tap*cin_pad + c
# ------------------------------------------------------------------------------------------------
def _pack_taps(w_tap_last: torch.Tensor, seg_sizes: Sequence[int]) -> Tuple[torch.Tensor, List[int], List[int]]:
    """w_tap_last: fp32 [rows, ntaps, cin_total].  Returns ([rows, K] fp32, kbase per seg, cin_pad per seg)."""
    rows, ntaps, _ = w_tap_last.shape
    blocks, kbase, cpads = [], [], []
    c0, k = 0, 0
    for cs in seg_sizes:
        cp = pad64(cs)
        blk = torch.zeros(rows, ntaps, cp, dtype=torch.float32, device=w_tap_last.device)
        blk[:, :, :cs] = w_tap_last[:, :, c0:c0 + cs]
        blocks.append(blk.reshape(rows, ntaps * cp))
        kbase.append(k)
        cpads.append(cp)
        k += ntaps * cp
        c0 += cs
    return torch.cat(blocks, dim=1), kbase, cpads


def taps_3x3():
    return [(0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]


def taps_3x3x3(pad_hw: int = 1):
    """pad_hw=1: 'same' conv; pad_hw=0: the stride-(1,2,2) down conv after F.pad(0,1,0,1,1,1)."""
    return [(kz - 1, ky - pad_hw, kx - pad_hw) for kz in range(3) for ky in range(3) for kx in range(3)]


@dataclass
class PackedWeight:
    w: torch.Tensor            # bf16 [rows_pad, ktot]  (fp32x: K = [hi blocks | lo blocks])
    kbase: List[int]           # per activation segment
    cin_pad: List[int]
    ktot: int
    rows: int                  # valid rows (nphase * cout)
    bias: Optional[torch.Tensor]
    taps: List[Tuple[int, int, int]]
    split: bool = False
    kbase_lo: Optional[List[int]] = None
    f16: bool = False          # `w` holds IEEE fp16 bit patterns (viewed as bfloat16 storage)


def pack_weight(w_rows_taps_c: torch.Tensor, seg_sizes: Sequence[int], taps, bias, device, split=False, row_mult=16,
                f16=False) -> PackedWeight:
    """w_rows_taps_c: fp32 [rows, ntaps, cin_total] (CPU or GPU).  f16: round to IEEE fp16 instead of bf16."""
    assert not (split and f16)
    w_rows_taps_c = w_rows_taps_c.to(device=device, dtype=torch.float32)
    mat, kbase, cpads = _pack_taps(w_rows_taps_c, seg_sizes)
    rows = mat.shape[0]
    rows_pad = (rows + row_mult - 1) // row_mult * row_mult
    if rows_pad != rows:
        mat = torch.cat([mat, torch.zeros(rows_pad - rows, mat.shape[1], device=device)], 0)
    k1 = mat.shape[1]
    kb_lo = None
    if split:
        hi, lo = split_hi_lo(mat)
        wq = torch.cat([hi, lo], dim=1).contiguous()
        kb_lo = [k + k1 for k in kbase]
    elif f16:
        wq = mat.to(torch.float16).contiguous().view(BF16)
    else:
        wq = mat.to(BF16).contiguous()
    b = None if bias is None else bias.to(device=device, dtype=torch.float32).contiguous()
    return PackedWeight(wq, kbase, cpads, wq.shape[1], rows, b, list(taps), split, kb_lo, bool(f16))


def pack_conv2d(w, seg_sizes, bias, device, split=False, f16=False):
    """nn.Conv2d weight [Cout, Cin, 3, 3] (unet/blocks.py:29-36)."""
    co, ci, kh, kw = w.shape
    assert (kh, kw) == (3, 3)
    return pack_weight(w.permute(0, 2, 3, 1).reshape(co, 9, ci), seg_sizes, taps_3x3(), bias, device, split, f16=f16)


def pack_conv3d(w, bias, device, split=False, down=False, f16=False):
    """nn.Conv3d weight [Cout, Cin, k, k, k], k in {1, 3} (vae/blocks.py:155-169, encoder.py:45,56)."""
    co, ci, kd, kh, kw = w.shape
    if kd == 1:
        return pack_weight(w.reshape(co, 1, ci), [ci], [(0, 0, 0)], bias, device, split, f16=f16)
    return pack_weight(w.permute(0, 2, 3, 4, 1).reshape(co, 27, ci), [ci], taps_3x3x3(0 if down else 1), bias, device, split, f16=f16)


def pack_conv3d_zstack(w, bias, device, f16=False):
    """Conv3d 3x3x3 with few input channels (encoder.py:30: 3, decoder.py:31: 8) over a z-stacked input (b2d_zstack_cl:
    channel kz*Cin + c holds slice z+kz-1): 9 in-plane taps on ONE 64-channel chunk instead of 27 taps on a chunk that is
    mostly zero padding -- a third of the MMAs."""
    co, ci = w.shape[:2]
    assert 3 * ci <= 64 and tuple(w.shape[2:]) == (3, 3, 3)
    rows = w.float().permute(0, 3, 4, 2, 1).reshape(co, 9, 3 * ci)   # [co][ky,kx][kz*ci + c]
    taps = [(0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    return pack_weight(rows, [3 * ci], taps, bias, device, f16=f16)


def pack_conv3d_zfold(w, device, f16=False):
    """Conv3d 3x3x3 with Cout <= 3 (decoder.py:71) as a per-slice 3x3 conv with rows (kz, co): row kz*4 + co holds
    w[co, :, kz] -- 9 taps and one N = 16 tile instead of 27 taps (an N = 16 MMA costs as much as an N = 64 one, so the
    tap count is what matters); b2d_zfold_combine sums the three z contributions and adds the bias."""
    co, ci = w.shape[:2]
    assert co <= 3 and tuple(w.shape[2:]) == (3, 3, 3)
    rows = torch.zeros(12, 9, ci, dtype=torch.float32)
    for kz in range(3):
        rows[kz * 4:kz * 4 + co] = w[:, :, kz].float().permute(0, 2, 3, 1).reshape(co, 9, ci)
    taps = [(0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    return pack_weight(rows, [ci], taps, None, device, f16=f16)


def pack_conv3d_upsampled(w, bias, device, py: int, px: int, split=False, f16=False):
    """Conv3d 3x3x3 applied to a nearest-2x (in-plane) upsampled map (decoder.py:46-47, 58-59), output phase (py, px):
    out[2y+py, 2x+px] reads only a 2x2 in-plane neighbourhood of the LOW-resolution map, with the 3x3 weights that land
    on the same source pixel summed (fp32) -- 12 taps instead of 27 on 4x the pixels (2.25x fewer MACs, no upsampled
    tensor in HBM).  Rows/cols: phase 0 reads offsets {-1: w0, 0: w1+w2}; phase 1 reads {0: w0+w1, +1: w2}."""
    co, ci = w.shape[:2]
    sets = {0: [(-1, (0,)), (0, (1, 2))], 1: [(0, (0, 1)), (1, (2,))]}
    taps, cols = [], []
    for kz in range(3):
        for dy, kys in sets[py]:
            for dx, kxs in sets[px]:
                taps.append((kz - 1, dy, dx))
                acc = torch.zeros(co, ci, dtype=torch.float32)
                for ky in kys:
                    for kx in kxs:
                        acc += w[:, :, kz, ky, kx].float()
                cols.append(acc)
    return pack_weight(torch.stack(cols, dim=1), [ci], taps, bias, device, split, f16=f16)


def pack_convT2x2(w, bias, device, split=False, f16=False):
    """nn.ConvTranspose2d k2 s2 weight [Cin, Cout, 2, 2] (unet/blocks.py:128-133): 4 phase GEMMs, rows phase-major."""
    ci, co, kh, kw = w.shape
    assert (kh, kw) == (2, 2)
    return pack_weight(w.permute(2, 3, 1, 0).reshape(4 * co, 1, ci), [ci], [(0, 0, 0)], bias, device, split, f16=f16)


def pack_linear(w, bias, device, split=False, f16=False):
    """nn.Linear / Conv1d-k1 weight [out, in]."""
    return pack_weight(w.reshape(w.shape[0], 1, -1), [w.reshape(w.shape[0], -1).shape[1]], [(0, 0, 0)], bias, device, split, f16=f16)


# ------------------------------------------------------------------------------------------------
# conv plans
# ------------------------------------------------------------------------------------------------
WORKSPACE_BYTES = 96 << 20
_DEFAULT_WS: dict = {}


def new_workspace(device, nbytes: int = WORKSPACE_BYTES) -> torch.Tensor:
    """A split-K scratch buffer (first 16 KB: arrival counters, zero-initialised once; every kernel leaves them zero).
    Plans that share one must run back to back on ONE stream: each predictor session / module program owns its own, so
    sampling loops on different streams never touch the same counters or partial tiles."""
    return torch.zeros(nbytes, dtype=torch.uint8, device=torch.device(device))


def default_workspace(device) -> torch.Tensor:
    """Scratch for plans created without an explicit workspace (unit tests, tools): one per (device, current stream)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    if key not in _DEFAULT_WS:
        _DEFAULT_WS[key] = new_workspace(torch.device("cuda", idx))
    return _DEFAULT_WS[key]


class ConvPlan:
    """Owns a b2d_conv_plan and keeps every tensor whose address is baked into it alive."""

    def __init__(self, inputs: Sequence[Act], pw: PackedWeight, out, *, cout: int, nphase: int = 1, stride: int = 1,
                 out_mode: int = 0, out_geom=None, residual: Optional[Act] = None, stats: Optional[torch.Tensor] = None,
                 stats_cpg: int = 0, out_scale=None, out_mask=None, block_n: int = 0, out_cstride=None, out_coff: int = 0,
                 in_norm=None, workspace: Optional[torch.Tensor] = None, tune_flags: int = 0, tune_ksplit: int = 0,
                 sched: Optional[dict] = None, in_temb=None):
        """in_norm = (stats, cpg, gamma, beta, act[, eps]): inputs[0] is a RAW pre-GroupNorm tensor and the kernel applies
        GroupNorm (+SiLU) to each staged tile (halo staging only; see include/b2d.h).
        in_temb = (table [T, ncols] fp32, row [N * stride] int32, row_stride, col): with in_norm, the DoubleBlock's time
        embedding table[row[n * stride], col + c] is added after the activation.
        workspace: split-K scratch (new_workspace); default: the (device, current stream) one.
        sched: out_mode 3 -- fuse the sampler update into the epilogue (b2d_conv_desc.sched_*): dict(kind, x, coef, state
        [int32: step index, ticket, 64-bit seed at word 2], noise=None, clip=(lo, hi) or None, x_bf16: Act or None,
        step_inc=1, philox=bool); `out` may then be None (eps is not stored)."""
        N, D, H, W, _ = inputs[0].shape
        d = ConvDesc()
        split = pw.split
        segs: List[Tuple[torch.Tensor, int, int]] = []  # (tensor, cin_pad, kbase)
        for i, a in enumerate(inputs):
            assert a.shape[:4] == (N, D, H, W) and a.C == pw.cin_pad[i], (a.shape, pw.cin_pad, i)
            assert in_norm is not None or a.f16 == pw.f16, "an MMA operand must be stored in the weights' 16-bit format"
            segs.append((a.hi, a.C, pw.kbase[i]))
        if split:
            for i, a in enumerate(inputs):
                assert a.lo is not None
                segs.append((a.hi, a.C, pw.kbase_lo[i]))   # A_hi * W_lo
                segs.append((a.lo, a.C, pw.kbase[i]))      # A_lo * W_hi
        assert len(segs) <= _lib.B2D_MAX_SEG
        d.nseg = len(segs)
        for i, (t, c, kb) in enumerate(segs):
            d.in_[i] = t.data_ptr()
            d.cin[i] = c
            d.kbase[i] = kb
        d.N, d.D, d.H, d.W = N, D, H, W
        d.ntaps = len(pw.taps)
        for i, (dz, dy, dx) in enumerate(pw.taps):
            d.tap_dz[i], d.tap_dy[i], d.tap_dx[i] = dz, dy, dx
        d.stride_h = d.stride_w = stride
        OH, OW = (H // stride, W // stride)
        d.OH, d.OW = OH, OW
        d.weight = pw.w.data_ptr()
        d.wrows, d.ktot = pw.w.shape
        d.cout, d.nphase = cout, nphase
        d.bias = ptr(pw.bias)
        out_hi = out.hi if isinstance(out, Act) else out
        d.out = ptr(out_hi)
        d.out_lo = ptr(out.lo) if isinstance(out, Act) else None
        d.out_f16 = 1 if (isinstance(out, Act) and out.f16) else 0
        d.out_mode = out_mode
        if out_geom is None:
            up = 2 if nphase == 4 else 1
            out_geom = (OH * up, OW * up, up, up, 0, 0)
        d.out_H, d.out_W, d.out_sy, d.out_sx, d.out_oy, d.out_ox = out_geom
        if out_cstride is None:
            out_cstride = cout if (out_mode == 1 or out_hi is None) else out_hi.shape[-1]
        d.out_cstride, d.out_coff = out_cstride, out_coff
        if residual is not None:
            d.residual = residual.hi.data_ptr()
            d.residual_lo = ptr(residual.lo)
            d.res_cstride = residual.C
            d.res_f16 = 1 if residual.f16 else 0
        d.stats = ptr(stats)
        d.stats_cpg = stats_cpg
        d.out_scale = ptr(out_scale)
        d.out_mask = ptr(out_mask)
        d.block_n = block_n
        d.tune_flags, d.tune_ksplit = tune_flags, tune_ksplit
        d.op_f16 = 1 if pw.f16 else 0
        if sched is not None:
            assert out_mode == 3
            st = sched["state"]
            assert st.dtype == torch.int32 and st.numel() >= 4 and st.data_ptr() % 8 == 0
            d.sched_kind = sched["kind"]
            d.sched_x = sched["x"].data_ptr()
            d.sched_noise = ptr(sched.get("noise"))
            d.sched_coef = sched["coef"].data_ptr()
            d.sched_step_idx = st.data_ptr()
            d.sched_ticket = st.data_ptr() + 4
            d.sched_seed_dev = (st.data_ptr() + 8) if sched.get("philox") else None
            d.sched_seed = 0
            d.sched_step_off = 0
            d.sched_step_inc = sched.get("step_inc", 1)
            clip = sched.get("clip")
            d.sched_clip = 0 if clip is None else 1
            d.sched_clip_lo, d.sched_clip_hi = (0.0, 0.0) if clip is None else (float(clip[0]), float(clip[1]))
            xb = sched.get("x_bf16")
            if xb is not None:
                assert xb.f16 == pw.f16
                d.sched_x_bf16, d.sched_x_bf16_lo, d.sched_bf16_stride = xb.hi.data_ptr(), ptr(xb.lo), xb.C
        if in_norm is not None:
            st_in, cpg_in, g_in, b_in, act_in = in_norm[:5]
            assert len(inputs) == 1 and not split
            d.in_stats = st_in.data_ptr()
            d.in_gamma, d.in_beta = ptr(g_in), ptr(b_in)
            d.in_cpg = cpg_in
            d.in_creal = g_in.numel() if g_in is not None else inputs[0].C
            d.in_f16 = 1 if inputs[0].f16 else 0
            d.in_act = 1 if act_in else 0
            d.in_eps = in_norm[5] if len(in_norm) > 5 else 1e-5
            if in_temb is not None:
                tt, trow, tstride, tcol = in_temb
                assert tt.dtype == torch.float32 and tt.is_contiguous() and trow.dtype == torch.int32
                d.in_temb, d.in_temb_row = tt.data_ptr(), trow.data_ptr()
                d.in_temb_row_stride, d.in_temb_ncols, d.in_temb_col = tstride, tt.shape[1], tcol
        ws = workspace if workspace is not None else default_workspace(inputs[0].hi.device)
        d.workspace = ws.data_ptr()
        d.workspace_bytes = ws.numel()
        self._keep = (inputs, pw, out, residual, stats, out_scale, out_mask, in_norm, ws, sched, in_temb)
        self.desc = d
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().b2d_conv_plan_create(C.byref(d), C.byref(self.handle)), "b2d_conv_plan_create")
        self.flops = 2.0 * N * D * OH * OW * cout * nphase * sum(len(pw.taps) * c for (_, c, _) in segs[:len(inputs)])

    def info(self):
        gm, gn, bn, kb = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().b2d_conv_plan_info(self.handle, C.byref(gm), C.byref(gn), C.byref(bn), C.byref(kb)))
        return dict(grid_m=gm.value, grid_n=gn.value, block_n=bn.value, kblocks=kb.value)

    def info2(self):
        out = (C.c_int32 * 8)()
        _lib.check(_lib.lib().b2d_conv_plan_info2(self.handle, out))
        return dict(zip(("engine", "halo", "ksplit", "units", "ctas", "block_n", "kgroups", "ws_kib"), list(out)))

    def run(self, stream: int):
        call("b2d_conv_run", self.handle, stream)

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().b2d_conv_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# elementwise launches
# ------------------------------------------------------------------------------------------------
def gn_apply(x: Act, y: Act, stats: torch.Tensor, cpg: int, gamma, beta, act: bool, stream: int, *, temb=None, temb_row=None,
             temb_row_stride=0, temb_col=0, stats_out=None, eps=1e-5):
    N, D, H, W, Cc = x.shape
    call("b2d_gn_apply", ptr(x.hi), ptr(x.lo), ptr(y.hi), ptr(y.lo), N, D * H * W, Cc, ptr(stats), cpg, ptr(gamma), ptr(beta),
         eps, 1 if act else 0, ptr(temb), ptr(temb_row), temb_row_stride, 0 if temb is None else temb.shape[1], temb_col,
         ptr(stats_out), 1 if x.f16 else 0, 1 if y.f16 else 0, stream)


def planar_to_cl(x: torch.Tensor, y: Act, N: int, C: int, P: int, coff: int, div_scale, stream: int):
    """planar fp32 [N][C][P] (optionally / div_scale[c]) -> channels coff.. of the channels-last Act y, in y's format."""
    call("b2d_planar_to_cl", x.data_ptr(), ptr(y.hi), ptr(y.lo), N, C, P, y.C, coff, ptr(div_scale), 1 if y.f16 else 0, stream)


def fused_gn_ok(x: Act, elems_per_sample: int) -> bool:
    """The per-sample fused GroupNorm kernels (b2d_gn_gn_apply, b2d_maxpool2x2_gn) apply: 16-bit modes, a sample of at most
    65536 elements, C/8 dividing 1024."""
    return x.lo is None and elems_per_sample <= 65536 and 1024 % (x.C // 8) == 0


def gn_gn_apply(x: Act, y1: Act, y2: Act, stats1: torch.Tensor, g1, b1, act1: bool, g2, b2, act2: bool, stream: int, eps=1e-5):
    N, D, H, W, Cc = x.shape
    call("b2d_gn_gn_apply", ptr(x.hi), 1 if x.f16 else 0, ptr(y1.hi), ptr(y2.hi), N, D * H * W, Cc, ptr(stats1), ptr(g1), ptr(b1), eps,
         1 if act1 else 0, ptr(g2), ptr(b2), eps, 1 if act2 else 0, 1 if y1.f16 else 0, stream)


def maxpool_gn(x: Act, y: Act, g, b, act: bool, stream: int, eps=1e-5):
    N, D, H, W, Cc = x.shape
    assert D == 1
    assert x.f16 == y.f16
    call("b2d_maxpool2x2_gn", ptr(x.hi), ptr(y.hi), N, H, W, Cc, ptr(g), ptr(b), eps, 1 if act else 0, 1 if x.f16 else 0, stream)


def maxpool_stats(x: Act, y: Act, stats: torch.Tensor, stream: int):
    N, D, H, W, Cc = x.shape
    assert D == 1
    assert x.f16 == y.f16
    call("b2d_maxpool2x2_stats", ptr(x.hi), ptr(x.lo), ptr(y.hi), ptr(y.lo), N, H, W, Cc, ptr(stats), 1 if x.f16 else 0, stream)


def upsample2x(x: Act, y: Act, stream: int):
    N, D, H, W, Cc = x.shape
    call("b2d_upsample2x_nearest", ptr(x.hi), ptr(y.hi), N * D, H, W, Cc, stream)
    if x.lo is not None:
        call("b2d_upsample2x_nearest", ptr(x.lo), ptr(y.lo), N * D, H, W, Cc, stream)


class Program:
    """A recorded sequence of kernel launches over static buffers (replayable, graph-capturable).  A step is a callable
    `fn(stream)`, or a list of them -- one per VARIANT: the same launch bound to different caller buffers (the chunks of a
    micro-batched VAE pass, whose first and last layers read / write a different slice of the big batch each time)."""

    def __init__(self):
        self.steps: List[Tuple[str, object]] = []
        self.flops = 0.0

    def add(self, name: str, fn):
        self.steps.append((name, fn))

    def run(self, stream: int, variant: int = 0):
        for _, fn in self.steps:
            if isinstance(fn, (list, tuple)):
                fn = fn[variant if len(fn) > 1 else 0]  # a one-entry list does not depend on the variant
            fn(stream)

    def __len__(self):
        return len(self.steps)
