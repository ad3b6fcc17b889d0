"""B200LatentDiffusionPredictor -- the sampling hot path of the reference's
LatentDiffusionPredictor (Diffusion_model/src/predictor.py:754-1023) rebuilt over libb2d:

    E2D encode (once)  ->  [UNet eps-prediction -> scheduler step] x steps  ->  D3D decode

Same public calls: `predict(img, velocity_2d, noise=None)` (DDPM, predictor.py:754-896) and
`predict_ddim(img, velocity_2d, num_steps=50, eta=0.0, noise=None)` (predictor.py:898-1023), same
tensor shapes and the same attributes (`.model`, `.scheduler`, `.vae`, `.normalizer`).

What differs from the reference's control flow (results are unaffected):
  * the latent shape is computed analytically (h = H/4, depth = num_slices, C = latent) instead of
    running E2D on a zeros tensor (predictor.py:765-774, 916-925);
  * the conditioning is written ONCE into channels 8..16 of a persistent channels-last UNet input
    buffer, and the scheduler kernel writes x_t into channels 0..7 -- no torch.cat per step
    (predictor.py:844, 982);
  * no `.item()` per step: the timestep is a device counter the scheduler kernel advances, so one
    timestep (UNet + scheduler, ~95 launches) is captured into a CUDA graph and replayed;
  * the Euclidean distance transform runs on the GPU (csrc/edt.cu) instead of SciPy on the host
    (predictor.py:1096-1116);
  * denormalisation and the mask multiply are the D3D `conv_out` epilogue (predictor.py:1005-1021);
  * the sampler update (p_sample / ddim_sample) runs inside final_conv's epilogue: eps never reaches HBM and the step
    costs one launch less (`fuse_scheduler`);
  * large batches are micro-batched: E2D and D3D run in chunks of `vae_chunk` samples over one reusable set of
    activation buffers, while the UNet loop runs over ALL B x num_slices slice-images per launch.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, checkpoint, engine, sharding
from .engine import Act, new_act
from .scheduler import B200Scheduler
from .unet import B200UNet
from .vae import B200DualVAE


class MaxNormalizerParams:
    """Holds `scale_factors` like the reference's MaxNormalizer (normalizer.py:22-58)."""

    def __init__(self, scale_factors, device):
        self.scale_factors = torch.tensor(list(scale_factors), dtype=torch.float32, device=device)


class B200LatentDiffusionPredictor:
    type = "latent-diffusion"

    def __init__(self, model_name="UNet", model_kwargs: Optional[dict] = None, distance_transform=True, *,
                 unet_state: Dict[str, torch.Tensor], vae_state: Dict[str, torch.Tensor], norm_factors: Sequence[float],
                 num_slices: int = 11, num_timesteps: int = 1000, precision: str = "f16", use_graph: bool = True,
                 fuse_scheduler: bool = True, vae_chunk: int = 8, vae_options: Optional[dict] = None, device="cuda"):
        if model_name != "UNet":
            raise ValueError("only the 'UNet' denoiser exists in the reference (predictor.py:136)")
        if not torch.cuda.is_available():
            raise RuntimeError("B200LatentDiffusionPredictor needs a CUDA device (sm_100a); there is no CPU fallback")
        if len(list(norm_factors)) != 3:
            raise ValueError(f"norm_factors must hold one scale per velocity component (3), got {list(norm_factors)}")
        if vae_chunk < 1:
            raise ValueError("vae_chunk must be >= 1")
        model_kwargs = dict(model_kwargs or {})
        model_kwargs.setdefault("time_embedding_dim", 64)  # predictor.py:317-318
        self.device = torch.device(device)
        self.num_slices, self.num_timesteps = num_slices, num_timesteps
        self.distance_transform = distance_transform
        self.precision = precision
        self.split = precision == "fp32x"
        self.f16 = precision == "f16"
        self.use_graph = use_graph
        self.fuse_scheduler = fuse_scheduler
        self.vae_chunk = int(vae_chunk)
        self.model = B200UNet(**model_kwargs, precision=precision, num_timesteps=num_timesteps, device=device)
        self.model.load_state_dict(unet_state)
        self.scheduler = B200Scheduler(num_timesteps=num_timesteps, device=device)
        lat = model_kwargs.get("out_channels", 4)
        self.latent_channels = lat
        self.vae = B200DualVAE(3, lat, precision=precision, device=device, options=vae_options)
        self.vae.load_state_dict(vae_state)
        self.vae_is_dual = True
        self.normalizer = {"input": MaxNormalizerParams([1], device), "output": MaxNormalizerParams(norm_factors, device)}
        self._session: Optional[dict] = None

    # ------------------------------------------------------------------------------ construction from the reference
    @classmethod
    def from_spec(cls, spec: "checkpoint.PredictorSpec", **kw) -> "B200LatentDiffusionPredictor":
        return cls("UNet", dict(spec.model_kwargs), spec.distance_transform, unet_state=spec.unet_state, vae_state=spec.vae_state,
                   norm_factors=spec.norm_factors, num_slices=spec.num_slices, num_timesteps=spec.num_timesteps, **kw)

    @classmethod
    def from_reference(cls, predictor, **kw) -> "B200LatentDiffusionPredictor":
        """Build from a constructed reference `LatentDiffusionPredictor` (its modules' own state_dicts, norm factors,
        num_slices, num_timesteps); the PyTorch modules stay untouched.  This is the body of INTEGRATION.md's `use_b200()`."""
        return cls.from_spec(checkpoint.spec_from_reference(predictor), **kw)

    @classmethod
    def from_directory(cls, folder: str, device: str = "cuda", **kw) -> "B200LatentDiffusionPredictor":
        """`Predictor.from_directory(folder, device)` (predictor.py:222-250) for this path: `log.json` + `model.pt` in
        `folder`, the frozen dual VAE from the `vae_encoder_path` / `vae_decoder_path` directories named in the log."""
        return cls.from_spec(checkpoint.load_directory(folder), device=device, **kw)

    # ------------------------------------------------------------------------------ session
    def _get_session(self, B, S, H, W) -> dict:
        key = (B, S, H, W)
        if self._session is not None and self._session["key"] == key:
            return self._session
        self._session = None
        dev, sp = self.device, self.split
        lat = self.latent_channels
        mult = 4 << len(self.model.features)
        if H % mult or W % mult:
            raise ValueError(f"in-plane size {H}x{W} must be a multiple of {mult} (latent /4, {len(self.model.features)} UNet poolings)")
        h, w = H // 4, W // 4
        N = B * S
        # micro-batching: the VAE passes run over `chunk` samples at a time against ONE set of activation buffers
        # (25 GB per 8 samples of 11x256x256); a ragged tail re-runs the last `chunk` samples (idempotent)
        chunk = min(B, self.vae_chunk)
        starts = sharding.chunk_starts(B, chunk)
        ses = dict(key=key, B=B, S=S, H=H, W=W, h=h, w=w, N=N, chunk=chunk, starts=starts)
        # everything two concurrent sampling loops must not share lives in the session: split-K scratch + arrival
        # counters, and the loop state {step index, ticket, 64-bit Philox seed} the kernels read and advance
        ses["workspace"] = engine.new_workspace(dev)
        ses["state"] = torch.zeros(8, dtype=torch.int32, device=dev)
        cin_pad = engine.pad64(self.model.in_channels)
        ses["unet_in"] = new_act(N, 1, h, w, cin_pad, dev, sp, zero=True, f16=self.f16)
        ses["x"] = torch.zeros(N, h, w, lat, dtype=torch.float32, device=dev)      # fp32 master latent, channels-last
        ses["eps"] = torch.zeros(N, h, w, lat, dtype=torch.float32, device=dev)
        ses["z"] = None                                                             # host-injected step noise (lazy)
        ses["img"] = torch.zeros(B, S, 1, H, W, dtype=torch.float32, device=dev)
        ses["v2d"] = torch.zeros(B, S, 3, H, W, dtype=torch.float32, device=dev)
        ses["edt"] = torch.zeros(2, N, H, W, dtype=torch.float32, device=dev)
        ses["feats"] = torch.zeros(N, h, w, dtype=torch.float32, device=dev)
        ses["out"] = torch.zeros(B, S, 3, H, W, dtype=torch.float32, device=dev)
        ui = ses["unet_in"]
        hi5 = ui.hi.view(B, S, h, w, cin_pad)
        lo5 = None if not sp else ui.lo.view(B, S, h, w, cin_pad)
        lat_views = [Act(hi5[c0:c0 + chunk], None if lo5 is None else lo5[c0:c0 + chunk], self.f16) for c0 in starts]
        # E2D: (chunk,S,3,H,W)/s -> NDHWC bf16 -> mu written straight into unet_in channels [lat, 2*lat)
        ses["e2d_in"] = new_act(chunk, S, H, W, 64, dev, sp, zero=True, f16=self.f16)
        ses["e2d"] = self.vae.build_encoder("encoder_2d", chunk, S, H, W, x_in=ses["e2d_in"], out=lat_views, out_mode=0, out_coff=lat,
                                            out_cout=lat, workspace=ses["workspace"])
        # D3D: reads the latent out of unet_in (conv_in's packed weight is zero beyond channel `lat`)
        scale = self.normalizer["output"].scale_factors
        ses["d3d"] = self.vae.build_decoder("decoder_3d", chunk, S, h, w, z_in=lat_views, out=[ses["out"][c0:c0 + chunk] for c0 in starts],
                                            out_scale=scale, out_mask=[ses["img"][c0:c0 + chunk] for c0 in starts],
                                            workspace=ses["workspace"])
        ses["unet"] = None
        ses["temb_key"] = None
        ses["final"] = {}
        ses["graph"] = None
        self._session = ses
        return ses

    def _bind_unet(self, ses, timesteps: List[int]):
        """(Re)build the UNet program against a time-embedding table reordered by sampling step."""
        key = tuple(timesteps)
        if ses["temb_key"] == key:
            return
        idx = torch.tensor(timesteps, dtype=torch.long, device=self.device)
        temb_steps = self.model.temb_table.index_select(0, idx).contiguous()
        ses["unet"] = None  # release the previous program's buffers first
        ses["final"] = {}
        ses["unet"] = self.model.build_program(ses["N"], ses["h"], ses["w"], x_in=ses["unet_in"], temb_row=ses["state"], temb_row_stride=0,
                                               temb_table=temb_steps, workspace=ses["workspace"], final=False)
        ses["temb_steps"] = temb_steps
        ses["temb_key"] = key
        ses["graph"] = None

    def _final(self, ses, kind, coef, clip_range, host_noise: bool, philox: bool, want_eps: bool):
        """final_conv of the step: fused with the sampler update (default), or the plain eps-producing launch."""
        fkey = (self.fuse_scheduler, kind, coef.data_ptr(), tuple(clip_range), host_noise, philox, want_eps)
        plan = ses["final"].get(fkey)
        if plan is None:
            if self.fuse_scheduler:
                sched = dict(kind=kind, x=ses["x"], coef=coef, state=ses["state"], noise=ses["z"] if host_noise else None,
                             clip=clip_range, x_bf16=ses["unet_in"], step_inc=1, philox=philox)
                plan = self.model.final_plan(ses["unet"], eps_out=ses["eps"] if want_eps else None, sched=sched)
            else:
                plan = self.model.final_plan(ses["unet"], eps_out=ses["eps"], eps_mode=2)
            ses["final"][fkey] = plan
        return plan

    # ------------------------------------------------------------------------------ stages
    def _conditioning(self, ses, img, velocity_2d, s):
        """predictor.py:927-962: normalise, E2D mu, EDT + bilinear features -> channels 8..16 of unet_in."""
        B, S, H, W, h, w, N = (ses[k] for k in ("B", "S", "H", "W", "h", "w", "N"))
        lat = self.latent_channels
        if img.shape[1] == 1 and S > 1:
            img = img.expand(B, S, 1, H, W)  # one mask for every slice (the broadcast of predictor.py:887-894)
        ses["img"].copy_(img.reshape(B, S, 1, H, W), non_blocking=True)
        ses["v2d"].copy_(velocity_2d, non_blocking=True)
        xi = ses["e2d_in"]
        chunk = ses["chunk"]
        scale = self.normalizer["output"].scale_factors
        for i, c0 in enumerate(ses["starts"]):
            engine.planar_to_cl(ses["v2d"][c0:c0 + chunk], xi, chunk * S, 3, H * W, 0, scale, s)
            ses["e2d"]["program"].run(s, variant=i)
        if self.distance_transform:
            _lib.call("b2d_edt2d", ses["img"].data_ptr(), ses["edt"].data_ptr(), N, H, W, s, launches=2)
            src = ses["edt"][0]
        else:
            src = ses["img"]
        _lib.call("b2d_bilinear_resize", src.data_ptr(), ses["feats"].data_ptr(), N, H, W, h, w, s)
        ui = ses["unet_in"]
        engine.planar_to_cl(ses["feats"], ui, N, 1, h * w, 2 * lat, None, s)

    def _set_latent(self, ses, noise, s):
        """x <- noise (N, C, h, w) planar; fp32 master is channels-last, bf16 copy goes to unet_in[..., :C]."""
        N, h, w = ses["N"], ses["h"], ses["w"]
        lat = self.latent_channels
        noise = noise.to(self.device, torch.float32).reshape(N, lat, h, w)
        ses["x"].copy_(noise.permute(0, 2, 3, 1))
        ui = ses["unet_in"]
        engine.planar_to_cl(ses["x"], ui, N * h * w, lat, 1, 0, None, s)

    def _new_seed(self, ses):
        """A fresh 64-bit Philox key per call, drawn from torch's default generator (so torch.manual_seed makes a run
        reproducible and consecutive calls draw different noise, like the reference's randn_like).  It lives in the
        session's device state: a captured graph reads it at replay time."""
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item()) | 1
        ses["state"][2:4].copy_(torch.tensor([seed], dtype=torch.int64).view(torch.int32))
        return seed

    def _run_unet(self, ses, s):
        ses["unet"]["program"].run(s)

    def _one_step(self, ses, kind, coef, clip_range, s, *, host_noise=False, philox=False, want_eps=False, advance=True):
        """One sampling step: UNet body, then final_conv + sampler update (one launch when fused)."""
        self._run_unet(ses, s)
        plan = self._final(ses, kind, coef, clip_range, host_noise, philox, want_eps)
        if self.fuse_scheduler:
            if not advance:
                raise RuntimeError("the fused step always advances the device step counter")
            plan.run(s)
            return
        plan.run(s)
        ui, x, st = ses["unet_in"], ses["x"], ses["state"]
        _lib.call("b2d_scheduler_step", kind, x.data_ptr(), ses["eps"].data_ptr(), _lib.ptr(ses["z"]) if host_noise else None, x.data_ptr(),
                  x.numel(), coef.data_ptr(), st.data_ptr(), 0, 1 if advance else 0, 1, float(clip_range[0]), float(clip_range[1]),
                  None if self.split else ui.hi.data_ptr(), self.latent_channels, ui.C, 0, (st.data_ptr() + 8) if philox else None,
                  st.data_ptr() + 4, 1 if self.f16 else 0, s)
        if self.split:
            N, h, w = ses["N"], ses["h"], ses["w"]
            engine.planar_to_cl(x, ui, N * h * w, self.latent_channels, 1, 0, None, s)

    def _run_loop(self, ses, kind, coef, n_steps, step_noise, clip_range, record, philox):
        s = _lib.stream_ptr()
        ses["state"][0:2].zero_()  # step index and ticket (an aborted run may have left the ticket mid-count)
        graphable = self.use_graph and step_noise is None and record is None
        n_launch = len(ses["unet"]["program"]) + (1 if self.fuse_scheduler else (3 if self.split else 2))
        if not graphable:
            host_noise = step_noise is not None
            if host_noise and ses["z"] is None:
                ses["z"] = torch.zeros_like(ses["x"])
            for i in range(n_steps):
                if host_noise:
                    ses["z"].copy_(step_noise[i].to(self.device, torch.float32).reshape(ses["N"], self.latent_channels, ses["h"], ses["w"])
                                   .permute(0, 2, 3, 1))
                x_before = ses["x"].clone() if record is not None else None
                self._one_step(ses, kind, coef, clip_range, s, host_noise=host_noise, philox=philox and not host_noise,
                               want_eps=record is not None)
                if record is not None:
                    record.append((x_before.permute(0, 3, 1, 2).contiguous(), ses["eps"].permute(0, 3, 1, 2).contiguous(),
                                   ses["x"].permute(0, 3, 1, 2).contiguous()))
            return
        gkey = (kind, coef.data_ptr(), tuple(clip_range), philox, self.fuse_scheduler)
        if ses["graph"] is None or ses["graph"][0] != gkey:
            # warm-up outside capture (lazy cudaFuncSetAttribute), then capture one timestep
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                saved = ses["x"].clone()
                saved_in = ses["unet_in"].hi.clone()
                saved_lo = ses["unet_in"].lo.clone() if self.split else None
                self._one_step(ses, kind, coef, clip_range, side.cuda_stream, philox=philox)
                ses["x"].copy_(saved)
                ses["unet_in"].hi.copy_(saved_in)
                if self.split:
                    ses["unet_in"].lo.copy_(saved_lo)
                ses["state"][0:2].zero_()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step(ses, kind, coef, clip_range, _lib.stream_ptr(), philox=philox)
            ses["graph"] = (gkey, g)
        g = ses["graph"][1]
        for _ in range(n_steps):
            g.replay()
        _lib.launch_count += n_steps * n_launch

    def _decode(self, ses, s):
        """predictor.py:993-1021."""
        for i in range(len(ses["starts"])):
            ses["d3d"]["program"].run(s, variant=i)
        return ses["out"].clone()

    # ------------------------------------------------------------------------------ public API
    def _check_inputs(self, img, velocity_2d):
        if img.dim() != 5 or velocity_2d.dim() != 5:
            raise ValueError("img must be (batch, num_slices, 1, H, W) and velocity_2d (batch, num_slices, 3, H, W)")
        B, S = velocity_2d.shape[0], velocity_2d.shape[1]
        if velocity_2d.shape[2] != 3 or img.shape[2] != 1 or img.shape[0] != B or img.shape[1] not in (S, 1) \
                or tuple(img.shape[3:]) != tuple(velocity_2d.shape[3:]):
            raise ValueError(f"shape mismatch: img {tuple(img.shape)} velocity_2d {tuple(velocity_2d.shape)}")
        return B, S, img.shape[3], img.shape[4]

    def _empty_fields(self, S, H, W):
        """An empty batch (a rank's shard when the batch is smaller than the world, sharding.predict_sharded): the
        reference's modules map (0, ...) to (0, ...); no kernel is launched."""
        return torch.empty(0, S, 3, H, W, dtype=torch.float32, device=self.device)

    def encode_target(self, velocity_3d: torch.Tensor, velocity_2d: Optional[torch.Tensor] = None) -> torch.Tensor:
        """predictor.py:1042-1085: 3D velocity target (B, S, 3, H, W) -> E3D latents (B, S, latent, H/4, W/4): permute to
        (B, 3, S, H, W), MaxNormalizer (fused into the layout pass), deterministic E3D encoding (mu), permute back.
        `velocity_2d` is unused, as in the reference."""
        if velocity_3d.dim() != 5 or velocity_3d.shape[2] != 3:
            raise ValueError(f"expected velocity_3d (batch, num_slices, 3, H, W), got {tuple(velocity_3d.shape)}")
        if not velocity_3d.is_cuda:
            raise RuntimeError("B200LatentDiffusionPredictor runs on a CUDA device only (no CPU fallback)")
        x = velocity_3d.permute(0, 2, 1, 3, 4).contiguous().float()
        mu, _ = self.vae.encode_3d_deterministic(x, div_scale=self.normalizer["output"].scale_factors)
        return mu.permute(0, 2, 1, 3, 4)

    def conditioning_latents(self, img: torch.Tensor, velocity_2d: torch.Tensor):
        """The conditioning the reference builds at the top of forward() / predict() (predictor.py:646-719 = :927-962):
        the frozen E2D encoder's mu of the normalised 2D velocity and the EDT features of the mask, bilinearly resized to
        the latent grid (the depth interpolation of :707-715 is the identity: the latent keeps the slice count).
        Returns fp32 (cond (N, latent, h, w), feats (N, 1, h, w)), N = batch * num_slices."""
        B, S, H, W = self._check_inputs(img, velocity_2d)
        ses = self._get_session(B, S, H, W)
        self._conditioning(ses, img.to(self.device), velocity_2d.to(self.device), _lib.stream_ptr())
        lat, ui = self.latent_channels, ses["unet_in"]
        raw = ui.hi[..., lat:2 * lat]
        cond = raw.view(torch.float16).float() if ui.f16 else raw.float()
        if ui.lo is not None:
            cond = cond + ui.lo[..., lat:2 * lat].float()
        cond = cond.reshape(ses["N"], ses["h"], ses["w"], lat).permute(0, 3, 1, 2).contiguous()
        return cond, ses["feats"].reshape(ses["N"], 1, ses["h"], ses["w"]).clone()

    def forward(self, img: torch.Tensor, velocity_2d: torch.Tensor, x_start: Optional[torch.Tensor] = None,
                noise: Optional[torch.Tensor] = None, *, t: Optional[torch.Tensor] = None):
        """predictor.py:636-751, the noise prediction of the training / validation loop (helper.py:277-320): conditioning
        as above, `t = randint(0, T, (N,))` (:736; `t=` injects it, like `noise=`), `x_t = q_sample(x_start, t, noise)`,
        `noise_pred = UNet(cat[x_t, cond, feats], t)`.  Returns (noise_pred, noise), both (N, latent, h, w).
        No autograd graph is attached: optimisation steps go through `train.LatentDiffusionTrainer`."""
        if x_start is None:
            raise ValueError("forward() requires x_start (target latents) for training. Use predict() for inference.")  # :750
        cond, feats = self.conditioning_latents(img, velocity_2d)
        N, lat, h, w = cond.shape
        x0 = x_start.to(self.device, torch.float32).reshape(N, lat, h, w)
        noise = torch.randn_like(x0) if noise is None else noise.to(self.device, torch.float32).reshape(N, lat, h, w)
        if t is None:
            t = torch.randint(0, self.num_timesteps, (N,), device=self.device).long()
        t = t.to(self.device).long()
        x_t = self.scheduler.q_sample(x0, t, noise)
        return self.model(torch.cat([x_t, cond, feats], dim=1), t), noise

    __call__ = forward

    def ddim_timesteps(self, num_steps: int) -> List[int]:
        """predictor.py:965: `torch.linspace(T-1, 0, num_steps, device=device).long()` -- built on the CUDA device like
        the reference's production path, so the integer rounding is the same kernel's."""
        return torch.linspace(self.num_timesteps - 1, 0, num_steps, device=self.device).long().tolist()

    def predict_ddim(self, img, velocity_2d, num_steps: int = 50, eta: float = 0.0, noise=None, *, step_noise=None, record=None):
        """predictor.py:898-1023."""
        B, S, H, W = self._check_inputs(img, velocity_2d)
        if B == 0:
            return self._empty_fields(S, H, W)
        ses = self._get_session(B, S, H, W)
        s = _lib.stream_ptr()
        timesteps = self.ddim_timesteps(num_steps)
        self._bind_unet(ses, timesteps)
        ckey = ("ddim", tuple(timesteps), float(eta))
        if ses.get("coef_key") != ckey:
            ses["coef"] = self.scheduler.ddim_coef_rows(timesteps, eta).to(self.device)
            ses["coef_key"] = ckey
            ses["graph"] = None
            ses["final"] = {}
        self._conditioning(ses, img.to(self.device), velocity_2d.to(self.device), s)
        if noise is None:
            noise = torch.randn(ses["N"], self.latent_channels, ses["h"], ses["w"], device=self.device)
        self._set_latent(ses, noise, s)
        # eta == 0: no row draws noise -> the kernels without the Philox generator
        philox = eta != 0.0 and step_noise is None
        if philox:
            self._new_seed(ses)
        self._run_loop(ses, 1, ses["coef"], num_steps, step_noise, (-30.0, 30.0), record, philox)
        return self._decode(ses, s)

    def predict(self, img, velocity_2d, noise=None, *, step_noise=None, record=None):
        """predictor.py:754-896 (multi-step branch; num_timesteps == 1 takes the one-shot branch :823-838)."""
        B, S, H, W = self._check_inputs(img, velocity_2d)
        if B == 0:
            return self._empty_fields(S, H, W)
        ses = self._get_session(B, S, H, W)
        s = _lib.stream_ptr()
        T = self.num_timesteps
        timesteps = list(reversed(range(T)))
        self._bind_unet(ses, timesteps)
        ckey = ("ddpm", T)
        if ses.get("coef_key") != ckey:
            if T == 1:
                # x0 = clamp((x - sqrt(1-abar) eps)/sqrt(abar)): {a, b, c1=1, c2=0, s=0}
                ab = self.scheduler._host["alphas_cumprod"][0]
                row = torch.stack([torch.sqrt(ab), torch.sqrt(1 - ab), torch.tensor(1.0), *([torch.tensor(0.0)] * 5)]).float()
                ses["coef"] = row.reshape(1, 8).to(self.device)
            else:
                ses["coef"] = self.scheduler.ddpm_coef_rows(timesteps).to(self.device)
            ses["coef_key"] = ckey
            ses["graph"] = None
            ses["final"] = {}
        self._conditioning(ses, img.to(self.device), velocity_2d.to(self.device), s)
        if noise is None:
            noise = torch.randn(ses["N"], self.latent_channels, ses["h"], ses["w"], device=self.device)
        self._set_latent(ses, noise, s)
        philox = step_noise is None and T > 1
        if philox:
            self._new_seed(ses)
        self._run_loop(ses, 0, ses["coef"], T, step_noise, (-30.0, 30.0), record, philox)
        return self._decode(ses, s)

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("B200LatentDiffusionPredictor runs on a CUDA device only (no CPU fallback)")
        return self
