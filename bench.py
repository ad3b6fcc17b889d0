#!/usr/bin/env python
"""Benchmark of the latent-diffusion sampling hot path (BASELINE.json metric:
"3D flow-field predictions/sec at 1/2/4/8 B200; UNet step ms; % TC/HBM peak").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

One "step" = one full `predict_ddim` (E2D encode -> 50 DDIM steps of the conditioned UNet -> D3D
decode) over this rank's batch of synthetic 256x256x11 microstructures, bf16, random-init weights of
the named architecture (no dataset/checkpoint is reachable offline).  Weak scaling: every rank owns
`--batch-per-gpu` samples; no collective on the sampling path, one gather of the decoded fields at
the end.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic work per sample (BASELINE.md section 2, 2*MAC): used for the roofline numerators
FLOP_E2D = 7.393e12
FLOP_D3D = 10.230e12
FLOP_UNET_STEP = 95.85e9
ELEMS_PER_SAMPLE = 360448  # latent elements per sample (11 x 8 x 64 x 64)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=8)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32x"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--slices", type=int, default=11)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")  # B200_PROFILING.md


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[6]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of predictor.predict_ddim on host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_sample(size: int, slices: int, ddim_steps: int, sample_slices: int = 3, unet_steps: int = 2):
    """Time a bounded sample of the reference's CPU path (its own op sequence, predictor.py:898-1023,
    including the shape-probe E2D pass on zeros, :916-925) and scale to one full prediction.
    Returns (predictions_per_s, detail dict)."""
    import torch
    from diffusion_model_project_b200 import synth
    from oracle import predictor as opred, unet as ounet, vae as ovae
    from oracle.scheduler import OracleScheduler, ddim_timesteps

    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
    S = min(sample_slices, slices)
    img, v2d = synth.synth_inputs(1, num_slices=S, size=size, seed=2024)
    noise = synth.synth_noise(1, num_slices=S, latent_size=size // 4, seed=42)
    t0 = time.perf_counter()
    ovae.encoder_forward(vsd, torch.zeros(1, 3, S, size, size), "encoder_2d.")           # predictor.py:916-925
    t_probe = time.perf_counter() - t0
    t0 = time.perf_counter()
    v_lat, feats = opred.conditioning(vsd, img, v2d, synth.NORM_FACTORS, use_edt=True)   # :927-962
    t_cond = time.perf_counter() - t0
    sch = OracleScheduler(1000)
    ts = ddim_timesteps(1000, ddim_steps)
    x = noise.reshape(v_lat.shape)
    for i in range(unet_steps + 1):  # the first step is untimed: it pays one-off allocator / primitive-creation costs
        if i == 1:
            t0 = time.perf_counter()
        tb = torch.full((x.shape[0],), ts[i], dtype=torch.long)
        eps = ounet.unet_forward(usd, torch.cat([x, v_lat, feats], 1), tb)                # :982-985
        x = sch.ddim_sample(eps, x, ts[i], ts[i + 1], 0.0, (-30.0, 30.0))                 # :988
    t_unet = (time.perf_counter() - t0) / unet_steps
    t0 = time.perf_counter()
    opred.decode(vsd, x, 1, img, synth.NORM_FACTORS)                                      # :993-1021
    t_dec = time.perf_counter() - t0
    scale = slices / S
    t_pred = (t_probe + t_cond + t_dec + ddim_steps * t_unet) * scale
    detail = dict(cores=cores, threads=torch.get_num_threads(), t_probe_e2d_s=t_probe, t_conditioning_s=t_cond, t_unet_step_s=t_unet,
                  t_decode_s=t_dec, sample=f"1 sample, {S} of {slices} slices at {size}x{size}, E2D probe + E2D + EDT + {unet_steps} (warm) of "
                  f"{ddim_steps} DDIM steps + D3D timed; scaled linearly to {slices} slices and {ddim_steps} steps")
    return 1.0 / t_pred, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals = []
    detail = None
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_sample(args.size, args.slices, args.ddim_steps)
        if i >= args.warmup:
            vals.append(v)
    val = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "3D flow-field predictions/sec", "value": val, "unit": "predictions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"predict_ddim DDIM-{args.ddim_steps} end to end (E2D -> UNet loop -> D3D), {args.size}x{args.size}x{args.slices}, "
                               "reference op sequence on host CPU (oracle port; /root/reference is not present on the GPU box)"},
        "cpu_baseline": {"value": val, "unit": "predictions/s", "cores": detail["cores"], "kind": "port", "sample": detail["sample"]},
        "e2e": {"value": val, "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "detail": {k: v for k, v in detail.items() if k != "sample"}, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from diffusion_model_project_b200 import _lib, synth, sharding
    from diffusion_model_project_b200.predictor import B200LatentDiffusionPredictor

    torch.set_grad_enabled(False)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    peaks = load_peaks()

    cpu_base = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        v, det = cpu_reference_sample(args.size, args.slices, args.ddim_steps)
        cpu_base = {"value": v, "unit": "predictions/s", "cores": det["cores"], "kind": "port", "sample": det["sample"]}

    B, S, H = args.batch_per_gpu, args.slices, args.size
    usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
    pred = B200LatentDiffusionPredictor("UNet", dict(synth.UNET_KWARGS), True, unet_state=usd, vae_state=vsd,
                                        norm_factors=synth.NORM_FACTORS, num_slices=S, num_timesteps=1000, precision=args.precision,
                                        use_graph=not args.no_graph, device=dev)
    del usd, vsd
    # rank-local slice of the synthetic global batch (no communication)
    img, v2d = synth.synth_inputs(B, num_slices=S, size=H, seed=2024 + rank)
    noise = synth.synth_noise(B, num_slices=S, latent_size=H // 4, seed=42 + rank * B)
    img_h, v2d_h, noise_h = img.pin_memory(), v2d.pin_memory(), noise.pin_memory()
    img_d, v2d_d, noise_d = img.to(dev), v2d.to(dev), noise.to(dev)
    out_h = torch.empty(B, S, 3, H, H, dtype=torch.float32).pin_memory()
    h2d_bytes = img_h.numel() * 4 + v2d_h.numel() * 4 + noise_h.numel() * 4
    d2h_bytes = out_h.numel() * 4

    def step_resident():
        out = pred.predict_ddim(img_d, v2d_d, num_steps=args.ddim_steps, eta=0.0, noise=noise_d)
        return sharding.gather_predictions(out, B * world, dst=0) if world > 1 else out

    def step_e2e():
        a = img_h.to(dev, non_blocking=True)
        b = v2d_h.to(dev, non_blocking=True)
        c = noise_h.to(dev, non_blocking=True)
        out = pred.predict_ddim(a, b, num_steps=args.ddim_steps, eta=0.0, noise=c)
        if world > 1:
            sharding.gather_predictions(out, B * world, dst=0)
        out_h.copy_(out, non_blocking=True)
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        sync_all()
        if sampler:
            sampler.start()
            torch.cuda.profiler.start()  # no-op unless run under `ncu --profile-from-start off` (profiles/ launch list)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        if sampler:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), _lib.launch_count - l0, clocks

    for _ in range(max(args.warmup, 3)):
        step_resident()
    ms, launches, clocks = timed(step_resident, args.steps, ClockSampler(local))
    for _ in range(2):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    ms_step = ms / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_val = B * world / (ms_e2e / args.steps / 1e3)

    # ---- stage breakdown + rooflines (rank 0, measured live with CUDA events on the launch stream) ----
    ses = pred._session
    s = _lib.stream_ptr()

    def time_launches(fn, reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    t_e2d = time_launches(lambda: ses["e2d"]["program"].run(s), 3)
    t_d3d = time_launches(lambda: ses["d3d"]["program"].run(s), 3)
    t_unet = time_launches(lambda: pred._run_unet(ses, s), 10)
    # dominant kernel: the tcgen05 implicit-GEMM conv engine (>= 85 % of the step's device time).  Its roofline
    # launch is the heaviest single launch of the step (D3D conv_up: 3x3x3 on the upsampled map), timed alone with
    # CUDA events on the launch stream -> burst peak.
    dom, dom_name = None, ""
    for name, fn in ses["d3d"]["program"].steps:
        plan = getattr(fn, "__self__", None)
        if plan is not None and hasattr(plan, "flops") and (dom is None or plan.flops > dom.flops):
            dom, dom_name = plan, name
    t_dom = time_launches(lambda: dom.run(s), 10)
    tc_ach = dom.flops / (t_dom * 1e-3) / 1e12
    di = dom.info2()
    dd = dom.desc
    dom_label = (f"conv_v{di['engine']}_kernel<BN={di['block_n']},halo={di['halo']}> (D3D {dom_name}: {sum(dd.cin[i] for i in range(1))}->{dd.cout} "
                 f"3x3x3 @ {B}x{dd.D}x{dd.H}x{dd.W}, timed alone: burst peak)")
    # DRAM bytes of that launch from the committed `ncu --set full` capture of the same shape (profiles/), if present
    traffic = None
    shape_key = f"3d {B} {dd.cin[0]} {dd.cout} {dd.H}"
    tpath = os.path.join(ROOT, "profiles", "r1_dominant_kernel_ncu.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("shape", "").startswith(shape_key):
            traffic = tj.get("dram_traffic_bytes")
    # scheduler kernel on >= 64 samples' worth of latent (369 MB > L2) for an HBM-bound number
    n_el = ELEMS_PER_SAMPLE * 64
    xs, es, zs = (torch.randn(n_el, device=dev) for _ in range(3))
    coef = pred.scheduler._ddpm_table
    t_sched = time_launches(lambda: _lib.call("b2d_scheduler_step", 0, xs.data_ptr(), es.data_ptr(), zs.data_ptr(), xs.data_ptr(),
                                              n_el, coef.data_ptr(), None, 500, 0, 1, -30.0, 30.0, None, 0, 0, 0, s), 20)
    hbm_ach = 16.0 * n_el / (t_sched * 1e-3) / 1e9
    del xs, es, zs

    if rank == 0:
        flop_step = B * (FLOP_E2D + FLOP_D3D + args.ddim_steps * FLOP_UNET_STEP) * (S / 11.0) * (H / 256.0) ** 2
        line = {
            "metric": "3D flow-field predictions/sec", "value": value, "unit": "predictions/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (fp32-class hi/lo split)", "data": "synthetic",
            "config": {"workload": f"predict_ddim DDIM-{args.ddim_steps} end to end (E2D encode -> conditioned-UNet loop -> D3D decode), "
                                   f"{B} samples/GPU of {S}x{H}x{H}, UNet in17/out8 k3 zeros-pad attn 3..2, CUDA-graphed timestep loop, "
                                   "random-init weights (BASELINE.json configs[2]/[3] shape, 8 samples per GPU)",
                       "global_batch": B * world, "parallelism": f"batch-sharded x{world}, final gather only",
                       "l2": "working set (>= 1.4 GB per activation tensor) far exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_val, "unit": "predictions/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": tc_ach, "peak": peaks["tc_burst"], "unit": "TFLOP/s", "frac": tc_ach / peaks["tc_burst"],
                         "traffic": traffic, "kernel": dom_label,
                         "peak_source": peaks["src"], "flops_per_launch": dom.flops, "ms_per_launch": t_dom},
            "roofline_scheduler": {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm"],
                                   "bytes_per_launch": 16.0 * n_el, "ms_per_launch": t_sched,
                                   "note": "DDPM step with host noise, 16 B/element, 64 samples (369 MB > L2)"},
            "stages": {"e2d_ms": t_e2d, "unet_step_ms": t_unet, "d3d_ms": t_d3d,
                       "e2d_tflops": B * FLOP_E2D * (S / 11.0) * (H / 256.0) ** 2 / (t_e2d * 1e-3) / 1e12,
                       "unet_tflops": B * FLOP_UNET_STEP * (S / 11.0) * (H / 256.0) ** 2 / (t_unet * 1e-3) / 1e12,
                       "d3d_tflops": B * FLOP_D3D * (S / 11.0) * (H / 256.0) ** 2 / (t_d3d * 1e-3) / 1e12,
                       "step_tflops": flop_step / (ms_step * 1e-3) / 1e12, "frac_of_sustained_peak": flop_step / (ms_step * 1e-3) / 1e12 / peaks["tc_sustained"]},
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
