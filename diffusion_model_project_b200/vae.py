"""B200DualVAE -- drop-in for the inference branches of the reference's DualBranchVAE
(VAE_model/src/dual_vae/model.py:32-243; Encoder vae/encoder.py:9-145, Decoder vae/decoder.py:10-151,
ResidualBlock vae/blocks.py:136-186): E2D `encode_2d_deterministic`, D3D `decode_3d`, plus
E3D `encode_3d_deterministic` (same engine).  All Conv3d layers run on the tcgen05 implicit-GEMM
engine in NDHWC bf16; GroupNorm(32) sums come out of the producing conv's epilogue.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import torch

from . import _lib, engine
from .engine import Act, ConvPlan, Program, new_act, pad64


class _Branch:
    """Packed weights of one Encoder or Decoder."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, kind: str, device, split: bool, f16: bool = False):
        self.kind = kind
        self.w: Dict[str, object] = {}
        g = lambda k: sd[prefix + k].detach().to("cpu", torch.float32)

        def conv(name, down=False):
            self.w[name] = engine.pack_conv3d(g(f"{name}.weight"), g(f"{name}.bias"), device, split, down=down, f16=f16)

        def norm(name):
            self.w[name] = (g(f"{name}.weight").to(device).contiguous(), g(f"{name}.bias").to(device).contiguous())

        def res(name):
            norm(f"{name}.norm1"); conv(f"{name}.conv1"); norm(f"{name}.norm2"); conv(f"{name}.conv2")
            if (prefix + f"{name}.residual_layer.weight") in sd:
                conv(f"{name}.residual_layer")

        conv("conv_in")
        if not split and g("conv_in.weight").shape[1] in (3, 8):
            self.w["conv_in.zstack"] = engine.pack_conv3d_zstack(g("conv_in.weight"), g("conv_in.bias"), device, f16=f16)
        for r in ("res1_1", "res1_2", "res2_1", "res2_2", "res3_1", "res3_2"):
            res(r)
        if kind == "encoder":
            conv("down1", down=True); conv("down2", down=True)
        else:
            conv("conv_up1"); conv("conv_up2")
            if not split:  # upsample folded into the conv: four phase convs on the low-resolution map
                for name in ("conv_up1", "conv_up2"):
                    for ph in range(4):
                        self.w[f"{name}.phase{ph}"] = engine.pack_conv3d_upsampled(g(f"{name}.weight"), g(f"{name}.bias"), device,
                                                                                  ph >> 1, ph & 1, f16=f16)
        norm("norm_out"); conv("conv_out")
        if kind == "decoder" and not split and g("conv_out.weight").shape[0] <= 3:
            self.w["conv_out.zfold"] = engine.pack_conv3d_zfold(g("conv_out.weight"), device, f16=f16)
            self.w["conv_out.bias"] = g("conv_out.bias").to(device=device, dtype=torch.float32).contiguous()
        self.cin = g("conv_in.weight").shape[1]
        self.cout = g("conv_out.weight").shape[0]


# Lowering choices of the VAE programs.  The defaults are the measured-best forms; tests switch them off one at a time to
# compare each fused form against the direct one.
DEFAULT_OPTIONS = dict(
    zstack=True,          # conv_in over a z-stacked input: 9 taps instead of 27 (bf16 mode)
    zfold=True,           # conv_out (C -> 3) as a per-slice 3x3 conv with rows (kz, co) + a gather over z
    upsample_fold=True,   # nn.Upsample + Conv3d as four phase convs on the low-resolution map
    norm_fusion=False,    # GroupNorm + SiLU applied by the consumer conv to its staged tiles (256-wide N tiles only)
)


def _as_list(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


class _Builder:
    """Records a launch program for one branch over static NDHWC buffers."""

    def __init__(self, B, device, split, workspace=None, options=None, f16=False):
        self.B, self.dev, self.split, self.f16 = B, device, split, f16
        self.ws = workspace if workspace is not None else engine.new_workspace(device)
        self.opt = dict(DEFAULT_OPTIONS, **(options or {}))
        self.prog = Program()
        self.keep: List[object] = []
        self.stats_buf = torch.zeros(B * 32 * 2 * 40, dtype=torch.float64, device=device)
        self.stats_used = 0

    def stats(self):
        n = self.B * 32 * 2
        v = self.stats_buf[self.stats_used:self.stats_used + n]
        self.stats_used += n
        assert self.stats_used <= self.stats_buf.numel()
        return v

    def conv(self, name, x: Act, pw, cout, *, stride=1, want_stats=True, residual=None, out=None, raw=False, **kw):
        """raw=True: the output is only ever read by a GroupNorm apply or as a residual (never as an MMA operand),
        so bf16 mode stores it as fp16 (finer rounding ahead of the normalisation, same 2 bytes).
        x / out / out_mask may be lists: one plan per chunk variant (same weights, same intermediate buffers)."""
        if isinstance(x, (list, tuple)) or isinstance(out, (list, tuple)) or isinstance(kw.get("out_mask"), (list, tuple)):
            assert not want_stats and residual is None
            nv = max(len(v) for v in (x, out, kw.get("out_mask")) if isinstance(v, (list, tuple)))
            xs, outs, masks = _as_list(x, nv), _as_list(out, nv), _as_list(kw.pop("out_mask", None), nv)
            plans = [ConvPlan([xs[i]], pw, outs[i], cout=cout, stride=stride, out_mask=masks[i], workspace=self.ws, **kw) for i in range(nv)]
            self.prog.flops += plans[0].flops
            self.prog.add(name, [pl.run for pl in plans])
            self.keep.append(plans)
            return outs, None
        N, D, H, W, _ = x.shape
        fmt = (raw or self.f16) and not self.split  # raw outputs are fp16 in both 16-bit modes
        if out is None:
            out = new_act(N, D, H // stride, W // stride, cout, self.dev, self.split, f16=fmt)
        elif isinstance(out, Act) and out.f16 != fmt:
            out = Act(out.hi, out.lo, fmt)
        st = self.stats() if want_stats else None
        plan = ConvPlan([x], pw, out, cout=cout, stride=stride, residual=residual, stats=st,
                        stats_cpg=(cout // 32) if want_stats else 0, workspace=self.ws, **kw)
        self.prog.flops += plan.flops
        self.prog.add(name, plan.run)
        self.keep.append(plan)
        return out, st

    def gn_silu(self, name, x: Act, st, gnw, inplace: bool):
        y = x.as_fmt(self.f16) if inplace else new_act(*x.shape, self.dev, self.split, f16=self.f16)
        g, b = gnw
        C = x.C
        self.prog.add(name, lambda s: engine.gn_apply(x, y, st, C // 32, g, b, True, s))
        return y

    def fusable(self, x: Act, cout: int) -> bool:
        """GroupNorm + SiLU can be applied by the consumer conv itself (persistent halo engine: stride-1 3x3x3,
        in-plane extent a multiple of 16, <= 512 input channels, bf16 mode).  Only worth it with 256-wide N tiles:
        the conv kernels sit at the shared-memory bandwidth roofline (tensor-core operand reads + TMA writes), and
        at BLOCK_N = 128 the in-place tile rewrite costs more than the HBM pass it saves; at 256 it is a wash
        (profiles/README.md), so the fused path is opt-in (options["norm_fusion"]) and the default is a separate pass."""
        _, _, H, W, C = x.shape
        return self.opt["norm_fusion"] and (not self.split) and H % 16 == 0 and W % 16 == 0 and C <= 512 and cout % 256 == 0

    def norm_conv(self, name, norm_name, x: Act, st_x, gnw, pw, cout, *, inplace_norm: bool, **kw):
        """conv(silu(GroupNorm32(x))): one launch when fusable (the conv normalises its staged input tiles in shared
        memory), else a GroupNorm-apply pass followed by the conv."""
        if self.fusable(x, cout):
            g, b = gnw
            return self.conv(name, x, pw, cout, in_norm=(st_x, x.C // 32, g, b, True), **kw)
        h = self.gn_silu(norm_name, x, st_x, gnw, inplace=inplace_norm)
        return self.conv(name, h, pw, cout, **kw)

    def res(self, w, name, x: Act, st_x, cin, cout, want_stats=True, raw_out=True):
        """vae/blocks.py:173-186.  raw_out=False when the block output feeds a conv directly (down / upsample+conv)."""
        r, st_r = self.norm_conv(f"{name}.conv1", f"{name}.norm1", x, st_x, w[f"{name}.norm1"], w[f"{name}.conv1"], cout,
                                 inplace_norm=False, raw=True)
        skip = x
        if f"{name}.residual_layer" in w:
            assert x.f16 == self.f16
            skip, _ = self.conv(f"{name}.residual_layer", x, w[f"{name}.residual_layer"], cout, want_stats=False, raw=True)
        return self.norm_conv(f"{name}.conv2", f"{name}.norm2", r, st_r, w[f"{name}.norm2"], w[f"{name}.conv2"], cout,
                              inplace_norm=True, residual=skip, want_stats=want_stats, raw=raw_out)

    def conv_in(self, w, x_in: Union[Act, Sequence[Act]], cin: int, cout: int):
        """First conv of a branch (encoder.py:30 / decoder.py:31).  With few input channels the three z taps are moved into
        the channel dimension first (b2d_zstack_cl + engine.pack_conv3d_zstack): a third of the MMAs.
        x_in may be a list (chunk variants: the same program reads a different slice of the caller's batch each run)."""
        xl = list(x_in) if isinstance(x_in, (list, tuple)) else [x_in]
        N, D, H, W, C = xl[0].shape
        if "conv_in.zstack" in w and self.opt["zstack"]:
            xs = new_act(N, D, H, W, C, self.dev, False, zero=True, f16=self.f16)
            self.keep.append((xs, xl))
            self.prog.add("conv_in.zstack", [lambda s, xi=xi: _lib.call("b2d_zstack_cl", _lib.ptr(xi.hi), _lib.ptr(xs.hi), N * D, D, H * W, cin, C, s)
                                             for xi in xl])
            return self.conv("conv_in", xs, w["conv_in.zstack"], cout, raw=True)
        if len(xl) == 1:
            return self.conv("conv_in", xl[0], w["conv_in"], cout, raw=True)
        # chunk variants without the z-stack pass (fp32x mode): one plan per chunk input, one shared output + statistics
        out = new_act(N, D, H, W, cout, self.dev, self.split, f16=True)
        st = self.stats()
        plans = [ConvPlan([xi], w["conv_in"], out, cout=cout, stats=st, stats_cpg=cout // 32, workspace=self.ws) for xi in xl]
        self.prog.flops += plans[0].flops
        self.prog.add("conv_in", [pl.run for pl in plans])
        self.keep.append(plans)
        return out, st

    def finish(self):
        used, buf = self.stats_used, self.stats_buf
        self.prog.steps = [("stats.zero", lambda s: _lib.call("b2d_zero", buf.data_ptr(), used * 8, s))] + self.prog.steps
        return self.prog


class B200DualVAE:
    def __init__(self, in_channels: int = 3, latent_channels: int = 8, kernel_size: int = 3, share_encoders: bool = False,
                 share_decoders: bool = False, *, precision: str = "f16", device="cuda", options: Optional[dict] = None):
        if kernel_size != 3:
            raise NotImplementedError("B200DualVAE: kernel_size must be 3")
        if precision not in engine.PRECISIONS:
            raise ValueError(f"precision must be one of {engine.PRECISIONS}")
        self.in_channels, self.latent_channels = in_channels, latent_channels
        self.split = precision == "fp32x"
        self.f16 = precision == "f16"
        self.precision = precision
        self.device = torch.device(device)
        self.options = dict(DEFAULT_OPTIONS, **(options or {}))
        self.branches: Dict[str, _Branch] = {}
        self._cache: Dict[tuple, dict] = {}

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        """Reference DualBranchVAE keys (`encoder_2d.*`, `decoder_3d.*`, optionally `encoder_3d.*`);
        branches absent from `sd` are simply not available."""
        for b, kind in (("encoder_2d", "encoder"), ("encoder_3d", "encoder"), ("decoder_3d", "decoder"), ("decoder_2d", "decoder")):
            if any(k.startswith(b + ".") for k in sd):
                self.branches[b] = _Branch(sd, b + ".", kind, self.device, self.split, self.f16)
        self._cache.clear()
        return self

    # ------------------------------------------------------------------------------ programs
    def build_encoder(self, branch: str, B, D, H, W, *, x_in: Optional[Act] = None, out=None, out_mode=1, out_coff=0,
                      out_cout=None, workspace=None) -> dict:
        """encoder.py:83-145.  out: planar fp32 [B][D][2*latent][h][w] (mode 1) or a channels-last Act
        receiving the first `out_cout` channels (mu) at channel offset out_coff (mode 0); a LIST of outputs makes the
        program's last launch a chunk variant (program.run(stream, variant=i) writes out[i])."""
        if H % 4 or W % 4:
            raise ValueError("encoder input H, W must be divisible by 4")
        br = self.branches[branch]
        w = br.w
        bd = _Builder(B, self.device, self.split, workspace, self.options, self.f16)
        if x_in is None:
            x_in = new_act(B, D, H, W, pad64(br.cin), self.device, self.split, zero=True, f16=self.f16)
        x, st = bd.conv_in(w, x_in, br.cin, 128)
        x, st = bd.res(w, "res1_1", x, st, 128, 128)
        x, _ = bd.res(w, "res1_2", x, st, 128, 128, want_stats=False, raw_out=False)   # -> down1 (MMA operand)
        x, st = bd.conv("down1", x, w["down1"], 128, stride=2)                            # -> res2_1's 1x1x1 skip conv
        x, st = bd.res(w, "res2_1", x, st, 128, 256)
        x, _ = bd.res(w, "res2_2", x, st, 256, 256, want_stats=False, raw_out=False)   # -> down2
        x, st = bd.conv("down2", x, w["down2"], 256, stride=2)                            # -> res3_1's skip conv
        x, st = bd.res(w, "res3_1", x, st, 256, 512)
        x, st = bd.res(w, "res3_2", x, st, 512, 512)
        h, wd = H // 4, W // 4
        cout = br.cout if out_cout is None else out_cout
        if out is None:
            out = torch.empty((B, D, br.cout, h, wd), dtype=torch.float32, device=self.device)
        kw = dict(out_mode=out_mode, out_coff=out_coff)
        if out_mode == 1:
            kw["out_cstride"] = br.cout
        bd.norm_conv("conv_out", "norm_out", x, st, w["norm_out"], w["conv_out"], cout, inplace_norm=True, want_stats=False, out=out, **kw)
        return dict(program=bd.finish(), x_in=x_in, out=out, keep=bd.keep, stats=bd.stats_buf)

    def build_decoder(self, branch: str, B, D, h, w_, *, z_in=None, out=None, out_scale=None, out_mask=None, workspace=None) -> dict:
        """decoder.py:79-151.  out: planar fp32 [B][D][3][H][W] (optionally * out_scale[c] * out_mask).
        z_in / out / out_mask may be equally long lists: chunk variants of the first and last launches."""
        br = self.branches[branch]
        w = br.w
        bd = _Builder(B, self.device, self.split, workspace, self.options, self.f16)
        if z_in is None:
            z_in = new_act(B, D, h, w_, pad64(br.cin), self.device, self.split, zero=True, f16=self.f16)
        x, st = bd.conv_in(w, z_in, br.cin, 512)
        x, st = bd.res(w, "res1_1", x, st, 512, 512)
        x, _ = bd.res(w, "res1_2", x, st, 512, 512, want_stats=False, raw_out=False)   # -> upsample -> conv_up1
        for stage, (cin, cout, r1, r2, last) in enumerate(((512, 256, "res2_1", "res2_2", False), (256, 128, "res3_1", "res3_2", True)), 1):
            N_, D_, H_, W_, _ = x.shape
            if f"conv_up{stage}.phase0" in w and H_ % 16 == 0 and W_ % 16 == 0 and self.options["upsample_fold"]:
                # nn.Upsample(scale=(1,2,2)) + Conv3d == 4 phase convs with 2x2x3 taps on the low-resolution map
                y = new_act(N_, D_, 2 * H_, 2 * W_, cout, self.device, self.split, f16=True)
                st = bd.stats()
                for ph in range(4):
                    plan = ConvPlan([x], w[f"conv_up{stage}.phase{ph}"], y, cout=cout, stats=st, stats_cpg=cout // 32,
                                    out_geom=(2 * H_, 2 * W_, 2, 2, ph >> 1, ph & 1), workspace=bd.ws)
                    bd.prog.flops += plan.flops
                    bd.prog.add(f"conv_up{stage}.phase{ph}", plan.run)
                    bd.keep.append(plan)
                x = y
            else:
                up = new_act(N_, D_, 2 * H_, 2 * W_, cin, self.device, self.split, f16=self.f16)
                bd.prog.add(f"up{stage}", lambda s, x=x, up=up: engine.upsample2x(x, up, s))
                x, st = bd.conv(f"conv_up{stage}", up, w[f"conv_up{stage}"], cout, raw=True)
            x, st = bd.res(w, r1, x, st, cout, cout)
            x, st = bd.res(w, r2, x, st, cout, cout, want_stats=last, raw_out=last)     # stage 1 -> upsample -> conv_up2
        H, W = 4 * h, 4 * w_
        if out is None:
            out = torch.empty((B, D, br.cout, H, W), dtype=torch.float32, device=self.device)
        if "conv_out.zfold" in w and H % 16 == 0 and W % 16 == 0 and self.options["zfold"]:
            # Conv3d 128 -> 3: nine in-plane taps per slice with rows (kz, co), then a gather over z (engine.pack_conv3d_zfold)
            hn = bd.gn_silu("norm_out", x, st, w["norm_out"], inplace=True)
            P = torch.empty((B, D, H, W, 12), dtype=torch.float32, device=self.device)
            plan = ConvPlan([hn], w["conv_out.zfold"], P, cout=12, out_mode=2, out_cstride=12, workspace=bd.ws)
            bd.prog.flops += plan.flops
            bd.prog.add("conv_out.zfold", plan.run)
            bd.keep.append((plan, P))
            bias, co = w["conv_out.bias"], br.cout
            nv = len(out) if isinstance(out, (list, tuple)) else 1
            outs, masks = _as_list(out, nv), _as_list(out_mask, nv)
            bd.keep.append((outs, masks))
            bd.prog.add("conv_out.combine", [lambda s, o=o, m=m: _lib.call("b2d_zfold_combine", P.data_ptr(), B * D, D, H, W, co, _lib.ptr(bias),
                                                                            _lib.ptr(out_scale), _lib.ptr(m), o.data_ptr(), co, 0, s)
                                             for o, m in zip(outs, masks)])
        else:
            bd.norm_conv("conv_out", "norm_out", x, st, w["norm_out"], w["conv_out"], br.cout, inplace_norm=True, want_stats=False,
                         out=out, out_mode=1, out_cstride=br.cout, out_scale=out_scale, out_mask=out_mask)
        return dict(program=bd.finish(), z_in=z_in, out=out, keep=bd.keep, stats=bd.stats_buf)

    # ------------------------------------------------------------------------------ module API
    def _encode(self, branch: str, x: torch.Tensor, div_scale: Optional[torch.Tensor] = None):
        """div_scale (fp32 [C] on the device): divide channel c by div_scale[c] in the layout pass (MaxNormalizer fused)."""
        if branch not in self.branches:
            raise RuntimeError(f"B200DualVAE: no weights loaded for {branch}")
        if not x.is_cuda:
            raise RuntimeError("B200DualVAE runs on a CUDA device only (no CPU fallback)")
        B, C, D, H, W = x.shape
        key = (branch, B, D, H, W)
        st = self._cache.get(key)
        if st is None:
            st = self.build_encoder(branch, B, D, H, W)
            self._cache = {key: st}
        s = _lib.stream_ptr()
        xi = st["x_in"]
        x = x.contiguous().float()
        engine.planar_to_cl(x, xi, B, C, D * H * W, 0, div_scale, s)
        st["program"].run(s)
        o = st["out"].permute(0, 2, 1, 3, 4).contiguous()  # [B][D][16][h][w] -> (B,16,D,h,w)
        mu, logvar = torch.chunk(o, 2, dim=1)
        return mu, logvar

    def encoder_2d(self, x):
        """Encoder.forward -> (mu, logvar)  (vae/encoder.py:83-145)."""
        return self._encode("encoder_2d", x)

    def encode_2d_deterministic(self, x):
        """dual_vae/model.py:225-233."""
        mu, logvar = self._encode("encoder_2d", x)
        return mu, (mu, torch.clamp(logvar, -10.0, 10.0))

    def encode_3d_deterministic(self, x, *, div_scale: Optional[torch.Tensor] = None):
        """dual_vae/model.py:235-243.  div_scale: optional per-channel divisor applied to x on the way in."""
        mu, logvar = self._encode("encoder_3d", x, div_scale)
        return mu, (mu, torch.clamp(logvar, -10.0, 10.0))

    def decode_3d(self, z: torch.Tensor) -> torch.Tensor:
        """dual_vae/model.py:211-223: (B, latent, D, h, w) -> (B, 3, D, 4h, 4w)."""
        if "decoder_3d" not in self.branches:
            raise RuntimeError("B200DualVAE: no weights loaded for decoder_3d")
        if not z.is_cuda:
            raise RuntimeError("B200DualVAE runs on a CUDA device only (no CPU fallback)")
        B, C, D, h, w = z.shape
        key = ("decoder_3d", B, D, h, w)
        st = self._cache.get(key)
        if st is None:
            st = self.build_decoder("decoder_3d", B, D, h, w)
            self._cache = {key: st}
        s = _lib.stream_ptr()
        zi = st["z_in"]
        z = z.contiguous().float()
        engine.planar_to_cl(z, zi, B, C, D * h * w, 0, None, s)
        st["program"].run(s)
        return st["out"].permute(0, 2, 1, 3, 4).contiguous()

    def eval(self):
        return self

    def parameters(self):
        return iter(())
