"""B200UNet -- drop-in for the reference's conditioned UNet
(Diffusion_model/src/unet/models.py:29-188, blocks.py:6-235): same constructor kwargs, same
state-dict keys, same `forward(x: (N,17,h,w) fp32 NCHW, time: (N,) long) -> (N,8,h,w)`.

Every per-step operation is a libb2d (sm_100a) kernel: 3x3 convs, the k2s2 transposed convs and
the attention projections run on the tcgen05 implicit-GEMM engine; GroupNorm statistics come out
of the producing kernel's epilogue and are applied (+SiLU, +time embedding) by one elementwise
pass; torch.cat is two K-segments of the consumer conv.  The time-embedding MLPs depend only on
(weights, t) and are precomputed into a [num_timesteps, sum(mid_channels)] table when weights
are loaded (models.py:14-26,72-82, blocks.py:89-95).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, engine
from .engine import Act, ConvPlan, Program, new_act, pad64
from .synth import attention_heads


class B200UNet:
    def __init__(self, in_channels=9, out_channels=4, features: Sequence[int] = (64, 128, 256, 512), kernel_size=3,
                 padding_mode="reflect", activation="silu", final_activation=None, attention: str = "", dropout: float = 0.0,
                 time_embedding_dim: Optional[int] = None, *, precision: str = "f16", num_timesteps: int = 1000,
                 device="cuda"):
        # same argument validation surface as the reference for the configurations this path supports
        if kernel_size != 3:
            raise NotImplementedError("B200UNet: kernel_size must be 3 (the shipped model)")
        if padding_mode != "zeros":
            raise NotImplementedError("B200UNet: padding_mode must be 'zeros' (config.py:256-261, the shipped model)")
        if (activation or "").strip().lower() != "silu" or final_activation:
            raise NotImplementedError("B200UNet: activation must be 'silu' and final_activation None")
        if dropout != 0.0:
            raise NotImplementedError("B200UNet: inference path, dropout must be 0")
        if precision not in engine.PRECISIONS:
            raise ValueError(f"precision must be one of {engine.PRECISIONS}")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.features = list(features)
        self.attention = attention
        self._heads = attention_heads(attention, len(self.features))
        self.time_embedding_dim = time_embedding_dim
        self.precision = precision
        self.split = precision == "fp32x"
        self.f16 = precision == "f16"   # 16-bit operand format: IEEE fp16 ("f16") or bf16 ("bf16"; fp32x: bf16 hi + lo)
        self.num_timesteps = num_timesteps
        self.device = torch.device(device)
        self._w: Dict[str, object] = {}
        self._programs: Dict[tuple, dict] = {}
        self.temb_table: Optional[torch.Tensor] = None
        self._temb_cols: Dict[str, int] = {}
        self.conv_tune_flags = 0   # _lib.TUNE_* bits handed to every conv plan (A/B measurements; 0 in production)
        # block2's halo-staged conv applies block1's GroupNorm + SiLU + time embedding to its staged tiles (16-bit modes,
        # maps that are multiples of 16 x 16, <= 512 channels) instead of a separate pass.  Parity-green, measured SLOWER
        # (UNet step 1841 -> 2006 us at 88 slice-images, 9689 -> 10417 at 704: the six convs lose more to the tile rewrite
        # than the six norm passes cost; profiles/README.md): off
        self.fuse_block1_norm = False

    # ------------------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """Consumes the reference UNet's state dict (keys of unet/models.py, unet/blocks.py) and
        repacks it once; the given tensors are never modified."""
        dev, sp, f16 = self.device, self.split, self.f16
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
        f = self.features
        w = {}

        def gn(prefix):
            return (sd[f"{prefix}.weight"].to(dev).contiguous(), sd[f"{prefix}.bias"].to(dev).contiguous())

        def double(prefix, seg_sizes):
            w[f"{prefix}.block1.conv"] = engine.pack_conv2d(sd[f"{prefix}.block1.conv.weight"], seg_sizes, None, dev, sp, f16)
            w[f"{prefix}.block1.norm"] = gn(f"{prefix}.block1.norm")
            cmid = sd[f"{prefix}.block1.conv.weight"].shape[0]
            w[f"{prefix}.block2.conv"] = engine.pack_conv2d(sd[f"{prefix}.block2.conv.weight"], [cmid], None, dev, sp, f16)
            w[f"{prefix}.block2.norm"] = gn(f"{prefix}.block2.norm")

        def attn(prefix, c):
            w[f"{prefix}.norm"] = gn(f"{prefix}.norm")
            w[f"{prefix}.in_proj"] = engine.pack_linear(sd[f"{prefix}.mha.in_proj_weight"], sd[f"{prefix}.mha.in_proj_bias"], dev, sp, f16)
            # proj_out(out_proj(a)) = (Wp Wo) a + (Wp bo + bp): fold the two C x C projections offline (fp64)
            wo, bo = sd[f"{prefix}.mha.out_proj.weight"].double(), sd[f"{prefix}.mha.out_proj.bias"].double()
            wp, bp = sd[f"{prefix}.proj_out.weight"][:, :, 0].double(), sd[f"{prefix}.proj_out.bias"].double()
            w[f"{prefix}.out"] = engine.pack_linear((wp @ wo).float(), (wp @ bo + bp).float(), dev, sp, f16)

        cin = self.in_channels
        for lvl, c in enumerate(f):
            double(f"encoder.{lvl}.0", [cin])
            if self._heads[lvl] is not None:
                attn(f"encoder.{lvl}.1", c)
            w[f"encoder.{lvl}.2.norm"] = gn(f"encoder.{lvl}.2.norm")
            cin = c
        double("bottleneck", [f[-1]])
        rheads = list(reversed(self._heads))
        for lvl, c in enumerate(reversed(f)):
            w[f"decoder.{lvl}.0.conv"] = engine.pack_convT2x2(sd[f"decoder.{lvl}.0.conv.weight"], sd[f"decoder.{lvl}.0.conv.bias"], dev, sp, f16)
            w[f"decoder.{lvl}.0.norm"] = gn(f"decoder.{lvl}.0.norm")
            double(f"decoder.{lvl}.1", [c, c])
            if rheads[lvl] is not None:
                attn(f"decoder.{lvl}.2", c)
        w["final_conv"] = engine.pack_conv2d(sd["final_conv.weight"], [f[0]], sd["final_conv.bias"], dev, sp, f16)
        self._w = w
        self._build_time_table(sd)
        self._programs.clear()
        return self

    def _build_time_table(self, sd):
        """[num_timesteps, sum(mid)] fp32: time_mlp(sinusoid(t)) pushed through every block's SiLU->Linear.
        One-off weight preprocessing in fp32 (models.py:14-26,78-82,141-142; blocks.py:92-95,100-103)."""
        if self.time_embedding_dim is None:
            self.temb_table = None
            return
        dev = self.device
        dim = self.time_embedding_dim
        t = torch.arange(self.num_timesteps, dtype=torch.float32)
        half = dim // 2
        e = math.log(10000) / (half - 1)
        e = torch.exp(torch.arange(half) * -e)
        e = t[:, None] * e[None, :]
        emb = torch.cat((e.sin(), e.cos()), dim=-1)
        emb = torch.nn.functional.linear(emb, sd["time_mlp.0.weight"], sd["time_mlp.0.bias"])
        emb = torch.nn.functional.linear(torch.nn.functional.silu(emb), sd["time_mlp.2.weight"], sd["time_mlp.2.bias"])
        act = torch.nn.functional.silu(emb)
        cols, parts, off = {}, [], 0
        names = [f"encoder.{l}.0" for l in range(len(self.features))] + ["bottleneck"] + [f"decoder.{l}.1" for l in range(len(self.features))]
        for p in names:
            wt, bt = sd[f"{p}.time_mlp.1.weight"], sd[f"{p}.time_mlp.1.bias"]
            parts.append(torch.nn.functional.linear(act, wt, bt))
            cols[p] = off
            off += wt.shape[0]
        self.temb_table = torch.cat(parts, dim=1).contiguous().to(dev)
        self._temb_cols = cols

    # ------------------------------------------------------------------------------ program
    def build_program(self, N: int, h: int, w: int, *, x_in: Optional[Act] = None, eps_out: Optional[torch.Tensor] = None,
                      eps_mode: int = 1, temb_row: Optional[torch.Tensor] = None, temb_row_stride: int = 1,
                      temb_table: Optional[torch.Tensor] = None, fuse_small: bool = True, workspace: Optional[torch.Tensor] = None,
                      final: bool = True) -> dict:
        """Allocate static buffers and conv plans for a batch of N (h x w) maps and record the launch list.

        fuse_small: where a sample has at most 65536 elements, the last norm of a DoubleBlock + the attention pre-norm,
        and max-pool + norm, each run as ONE per-sample launch (engine.gn_gn_apply, engine.maxpool_gn) instead of two
        launches that meet through global statistics.
        workspace: split-K scratch of this program (engine.new_workspace; allocated if None).
        final=False: stop before final_conv (unet/models.py:185); the caller appends it with `final_plan` -- the sampling
        loop fuses the DDPM / DDIM update into that launch.

        x_in: channels-last input [N,1,h,w,pad64(in_channels)] (allocated if None);
        eps_out: fp32 output, planar [N,out,h,w] (eps_mode 1) or channels-last [N,h,w,out] (eps_mode 2).
        """
        if not self._w:
            raise RuntimeError("B200UNet: load_state_dict() must be called before forward()")
        dev, sp, f, F16 = self.device, self.split, self.features, self.f16
        ws = workspace if workspace is not None else engine.new_workspace(dev)
        nl = len(f)
        if h % (1 << nl) or w % (1 << nl):
            raise ValueError(f"B200UNet: input {h}x{w} must be divisible by {1 << nl} (five 2x2 max-pools)")
        W_ = self._w
        if temb_table is None:
            temb_table = self.temb_table  # bound now: the launch list must not see later changes
        prog = Program()
        stats_slots: List[tuple] = []  # (offset, size) in doubles
        stats_total = [0]

        def stats_alloc(n_groups):
            off = stats_total[0]
            stats_total[0] += N * n_groups * 2
            stats_slots.append((off, N * n_groups * 2))
            return off

        # first pass collects stats sizes lazily: allocate a generous buffer up front
        n_gn = 2 * (2 * nl + 1) + 2 * nl + 2 * sum(1 for x in self._heads if x is not None)
        stats_buf = torch.zeros(N * 2 * (n_gn + 4), dtype=torch.float64, device=dev)

        def stats_view(off, size):
            return stats_buf[off:off + size]

        if x_in is None:
            x_in = new_act(N, 1, h, w, pad64(self.in_channels), dev, sp, zero=True, f16=F16)
        if temb_row is None:
            temb_row = torch.zeros(N, dtype=torch.int32, device=dev)
            temb_row_stride = 1
        if eps_out is None and final:
            eps_out = torch.empty((N, self.out_channels, h, w) if eps_mode == 1 else (N, h, w, self.out_channels),
                                  dtype=torch.float32, device=dev)
        keep = []

        def conv_gn_act(name, inputs, pw, cout, H, Wd, gnw, *, temb_col=None, stats_out=None, nphase=1, second=None, in_norm=None,
                        in_temb=None, defer_norm=False):
            """conv (+GN sums in the epilogue) -> in-place GN apply + SiLU (+temb).  second = (name2, y2, gamma2, beta2): the
            same launch also writes y2 = GN(out; gamma2, beta2) (the attention pre-norm), per-sample fused form."""
            up = 2 if nphase == 4 else 1
            # raw conv output: fp16 storage in both 16-bit modes (8x finer than bf16 ahead of the normalisation), rewritten
            # in place in the operand format by the GroupNorm apply
            raw = new_act(N, 1, H * up, Wd * up, cout, dev, sp, f16=True)
            out = raw.as_fmt(F16)
            st = stats_view(stats_alloc(1), N * 2)
            plan = ConvPlan(inputs, pw, raw, cout=cout, nphase=nphase, stats=st, stats_cpg=cout, workspace=ws, tune_flags=self.conv_tune_flags,
                            in_norm=in_norm, in_temb=in_temb)
            prog.flops += plan.flops
            prog.add(f"{name}.conv", plan.run)
            g, b = gnw
            if defer_norm:   # the consumer applies this layer's GroupNorm + SiLU (+ time embedding) to its staged tiles
                keep.append(plan)
                return raw, st
            if second is not None:
                assert temb_col is None and stats_out is None
                name2, y2, g2, b2 = second
                prog.add(f"{name}.gn+{name2}", lambda s: engine.gn_gn_apply(raw, out, y2, st, g, b, True, g2, b2, False, s))
                keep.append(plan)
                return out
            prog.add(f"{name}.gn", lambda s, raw=raw, out=out, st=st, g=g, b=b, tc=temb_col, so=stats_out, cout=cout: engine.gn_apply(
                raw, out, st, cout, g, b, True, s, temb=temb_table if tc is not None else None,
                temb_row=temb_row if tc is not None else None, temb_row_stride=temb_row_stride, temb_col=tc or 0, stats_out=so))
            keep.append(plan)
            return out

        def double(prefix, inputs, cmid, cout, H, Wd, stats_out=None, second=None):
            tc = self._temb_cols.get(prefix) if temb_table is not None else None
            if self.fuse_block1_norm and not sp and H % 16 == 0 and Wd % 16 == 0 and cmid <= 512:
                # block2's conv is halo-staged: it normalises block1's raw output tile by tile in shared memory
                # (b2d_conv_desc.in_stats / in_temb) -- one launch and one pass over the activation less
                raw1, st1 = conv_gn_act(f"{prefix}.block1", inputs, W_[f"{prefix}.block1.conv"], cmid, H, Wd, W_[f"{prefix}.block1.norm"],
                                        defer_norm=True)
                g1, b1 = W_[f"{prefix}.block1.norm"]
                return conv_gn_act(f"{prefix}.block2", [raw1], W_[f"{prefix}.block2.conv"], cout, H, Wd, W_[f"{prefix}.block2.norm"],
                                   stats_out=stats_out, second=second, in_norm=(st1, cmid, g1, b1, True),
                                   in_temb=(temb_table, temb_row, temb_row_stride, tc) if tc is not None else None)
            a = conv_gn_act(f"{prefix}.block1", inputs, W_[f"{prefix}.block1.conv"], cmid, H, Wd, W_[f"{prefix}.block1.norm"], temb_col=tc)
            return conv_gn_act(f"{prefix}.block2", [a], W_[f"{prefix}.block2.conv"], cout, H, Wd, W_[f"{prefix}.block2.norm"],
                               stats_out=stats_out, second=second)

        def double_attention(prefix_d, prefix_a, inputs, c, H, Wd, heads):
            """DoubleBlock followed by the level's SelfAttention (if any); their two norms share a launch when fusable."""
            if heads is None:
                return double(prefix_d, inputs, c, c, H, Wd)
            if fuse_small and not sp and H * Wd * c <= 65536 and 1024 % (c // 8) == 0:
                xn = new_act(N, 1, H, Wd, c, dev, sp, f16=F16)
                g2, b2 = W_[f"{prefix_a}.norm"]
                x = double(prefix_d, inputs, c, c, H, Wd, second=(f"{prefix_a}.gn", xn, g2, b2))
                return attention(prefix_a, x, c, H, Wd, heads, None, xn=xn)
            st_attn = stats_view(stats_alloc(1), N * 2)
            x = double(prefix_d, inputs, c, c, H, Wd, stats_out=st_attn)
            return attention(prefix_a, x, c, H, Wd, heads, st_attn)

        def attention(prefix, x: Act, c, H, Wd, heads, st_in, xn=None):
            """x <- x + (Wp Wo) softmax(q k^T/sqrt(d)) v + b, q,k,v = in_proj(GN(x))  (blocks.py:209-235).
            xn given: the pre-norm GN(x) was already written by the producer's fused launch."""
            T = H * Wd
            g, b = W_[f"{prefix}.norm"]
            if xn is None:
                xn = new_act(N, 1, H, Wd, c, dev, sp, f16=F16)
                prog.add(f"{prefix}.gn", lambda s: engine.gn_apply(x, xn, st_in, c, g, b, False, s))
            qkv = new_act(N, 1, H, Wd, 3 * c, dev, sp, f16=F16)
            p1 = ConvPlan([xn], W_[f"{prefix}.in_proj"], qkv, cout=3 * c, workspace=ws, tune_flags=self.conv_tune_flags)
            prog.add(f"{prefix}.in_proj", p1.run)
            ao = new_act(N, 1, H, Wd, c, dev, sp, f16=F16)
            prog.add(f"{prefix}.core", lambda s: _lib.call("b2d_attention", _lib.ptr(qkv.hi), _lib.ptr(qkv.lo), _lib.ptr(ao.hi),
                                                            _lib.ptr(ao.lo), N, T, c, heads, 1 if F16 else 0, s))
            p2 = ConvPlan([ao], W_[f"{prefix}.out"], x, cout=c, residual=x, workspace=ws, tune_flags=self.conv_tune_flags)
            prog.add(f"{prefix}.out_proj", p2.run)
            prog.flops += p1.flops + p2.flops + 4.0 * N * T * T * c
            keep.extend([p1, p2, xn, qkv, ao])
            return x

        skips = []
        x = x_in
        H, Wd = h, w
        for lvl, c in enumerate(f):
            heads = self._heads[lvl]
            x = double_attention(f"encoder.{lvl}.0", f"encoder.{lvl}.1", [x], c, H, Wd, heads)
            skips.append(x)
            pooled = new_act(N, 1, H // 2, Wd // 2, c, dev, sp, f16=F16)
            g, b = W_[f"encoder.{lvl}.2.norm"]
            if fuse_small and engine.fused_gn_ok(x, (H // 2) * (Wd // 2) * c):
                prog.add(f"encoder.{lvl}.2.pool+gn", lambda s, x=x, pooled=pooled, g=g, b=b: engine.maxpool_gn(x, pooled, g, b, True, s))
                x = pooled
                H, Wd = H // 2, Wd // 2
                continue
            st = stats_view(stats_alloc(1), N * 2)
            prog.add(f"encoder.{lvl}.2.pool", lambda s, x=x, pooled=pooled, st=st: engine.maxpool_stats(x, pooled, st, s))
            prog.add(f"encoder.{lvl}.2.gn", lambda s, pooled=pooled, st=st, g=g, b=b, c=c: engine.gn_apply(pooled, pooled, st, c, g, b, True, s))
            x = pooled
            H, Wd = H // 2, Wd // 2
        x = double("bottleneck", [x], 2 * f[-1], 2 * f[-1], H, Wd)
        rheads = list(reversed(self._heads))
        for lvl, c in enumerate(reversed(f)):
            up = conv_gn_act(f"decoder.{lvl}.0", [x], W_[f"decoder.{lvl}.0.conv"], c, H, Wd, W_[f"decoder.{lvl}.0.norm"], nphase=4)
            H, Wd = H * 2, Wd * 2
            heads = rheads[lvl]
            x = double_attention(f"decoder.{lvl}.1", f"decoder.{lvl}.2", [skips[nl - 1 - lvl], up], c, H, Wd, heads)
        assert stats_total[0] <= stats_buf.numel()
        used = stats_total[0]
        prog.steps = [("stats.zero", lambda s: _lib.call("b2d_zero", stats_buf.data_ptr(), used * 8, s))] + prog.steps
        st = dict(program=prog, x_in=x_in, eps=eps_out, temb_row=temb_row, keep=keep, stats=stats_buf, skips=skips,
                  temb_table=temb_table, final_in=x, workspace=ws, eps_mode=eps_mode)
        # algorithmic FLOPs of final_conv belong to the step whether or not the caller fuses the sampler update into it
        prog.flops += 2.0 * N * h * w * self.out_channels * 9 * x.C
        if final:
            pf = self.final_plan(st, eps_out=eps_out, eps_mode=eps_mode)
            prog.add("final_conv", pf.run)
            keep.append(pf)
        return st

    def final_plan(self, st: dict, *, eps_out: Optional[torch.Tensor] = None, eps_mode: int = 2, sched: Optional[dict] = None) -> ConvPlan:
        """final_conv (unet/models.py:185) over the program's last activation.  sched: fuse the DDPM / DDIM update that
        consumes eps into the epilogue (engine.ConvPlan `sched`; eps is stored only if eps_out is given, channels-last)."""
        if sched is not None:
            return ConvPlan([st["final_in"]], self._w["final_conv"], eps_out, cout=self.out_channels, out_mode=3,
                            out_cstride=self.out_channels, workspace=st["workspace"], sched=sched)
        return ConvPlan([st["final_in"]], self._w["final_conv"], eps_out, cout=self.out_channels, out_mode=eps_mode,
                        out_cstride=self.out_channels, workspace=st["workspace"])

    # ------------------------------------------------------------------------------ module API
    def forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None) -> torch.Tensor:
        """models.py:131-188.  x: (N, in_channels, h, w) float32 on the GPU; time: (N,) integer timesteps."""
        if self.time_embedding_dim is not None and time is None:
            raise ValueError("Model requires time input but None was provided")  # models.py:139-140
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected input (N,{self.in_channels},h,w), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("B200UNet runs on a CUDA device only (no CPU fallback)")
        N, _, h, w = x.shape
        if time is not None:
            time = torch.as_tensor(time)
            if time.numel() != N:
                raise ValueError(f"time must hold one timestep per image ({N}), got {tuple(time.shape)}")
            if self.temb_table is not None and (int(time.min()) < 0 or int(time.max()) >= self.num_timesteps):
                # the reference evaluates the sinusoid for any t (models.py:14-26); this path reads a table built for
                # t in [0, num_timesteps) -- refuse instead of reading outside it
                raise ValueError(f"timesteps must lie in [0, {self.num_timesteps}), got [{int(time.min())}, {int(time.max())}]")
        key = (N, h, w)
        st = self._programs.get(key)
        if st is None:
            st = self.build_program(N, h, w)
            self._programs = {key: st}  # one cached shape: buffers are large
        s = _lib.stream_ptr()
        x = x.contiguous().float()
        xi = st["x_in"]
        engine.planar_to_cl(x, xi, N, self.in_channels, h * w, 0, None, s)
        if time is not None:
            st["temb_row"].copy_(time.to(device=self.device, dtype=torch.int32).reshape(N), non_blocking=True)
        st["program"].run(s)
        return st["eps"].clone()

    __call__ = forward

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("B200UNet runs on a CUDA device only (no CPU fallback)")
        return self
