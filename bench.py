#!/usr/bin/env python
"""Benchmark of the latent-diffusion sampling hot path (BASELINE.json metric:
"3D flow-field predictions/sec at 1/2/4/8 B200; UNet step ms; % TC/HBM peak").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

One "step" = one full `predict_ddim` (E2D encode -> 50 DDIM steps of the conditioned UNet -> D3D
decode) over this rank's batch of synthetic 256x256x11 microstructures, IEEE-fp16 operands (`--precision`), random-init weights of
the named architecture (no dataset/checkpoint is reachable offline).  Default: weak scaling, every
rank owns `--batch-per-gpu` samples (BASELINE configs[2]: 64 samples over 8 GPUs = 8 per GPU); the
same line also carries `strong`: a FIXED global batch (`--strong-batch`, default 64 = configs[3], the
"test-set-sized batch") sharded B/N per rank, micro-batched through the VAE on each GPU.
`--global-batch G` makes that fixed batch the primary measurement (`"scaling": "strong"`).  No
collective on the sampling path, one gather of the decoded fields at the end.  Prints ONE JSON line
on rank 0.
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
    ap.add_argument("--global-batch", type=int, default=0, help="fixed global batch sharded over the ranks (strong scaling) as the primary measurement")
    ap.add_argument("--strong-batch", type=int, default=64, help="fixed global batch of the secondary `strong` object (0 = skip)")
    ap.add_argument("--vae-chunk", type=int, default=8)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16", "fp32x"])
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
        self.index, self.proc, self.lines, self.mark_idx = index, None, [], 0

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

    def mark(self):
        """The timed region starts here: the sampler has been running since before the warm-up steps (same workload), so
        a region shorter than nvidia-smi's start-up + sampling period still gets its samples."""
        self.mark_idx = len(self.lines)

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
        lines, window = self.lines[self.mark_idx:], "timed region"
        if not any(len(ln.split(",")) >= 7 for ln in lines) and self.mark_idx > 0:
            # a timed region shorter than one sampling period: the last samples of the warm-up steps just before it
            lines, window = self.lines[max(0, self.mark_idx - 3):self.mark_idx], "warm-up steps just before the timed region (region shorter than one 200 ms sample)"
        for ln in lines:
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
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window}


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of predictor.predict_ddim on host cores
# --------------------------------------------------------------------------------------------------
_CPU_CASE = {}


def cpu_reference_prediction(size: int, slices: int, ddim_steps: int):
    """ONE full prediction of the reference's CPU path: its own op sequence (predictor.py:898-1023, including the
    shape-probe E2D pass on zeros, :916-925) restated by the oracle, B = 1, all host cores, nothing scaled.
    Returns (seconds, detail dict)."""
    import torch
    from diffusion_model_project_b200 import synth
    from oracle import predictor as opred, vae as ovae

    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    key = (size, slices)
    if key not in _CPU_CASE:
        usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
        img, v2d = synth.synth_inputs(1, num_slices=slices, size=size, seed=2024)
        noise = synth.synth_noise(1, num_slices=slices, latent_size=size // 4, seed=42)
        _CPU_CASE[key] = (usd, vsd, img, v2d, noise)
    usd, vsd, img, v2d, noise = _CPU_CASE[key]
    t0 = time.perf_counter()
    ovae.encoder_forward(vsd, torch.zeros(1, 3, slices, size, size), "encoder_2d.")       # predictor.py:916-925
    t_probe = time.perf_counter() - t0
    out = opred.predict_ddim(usd, vsd, img, v2d, noise, num_steps=ddim_steps, eta=0.0, norm_factors=synth.NORM_FACTORS)
    t_all = time.perf_counter() - t0
    assert out.shape == (1, slices, 3, size, size)
    detail = dict(cores=cores, threads=torch.get_num_threads(), t_probe_e2d_s=t_probe, t_prediction_s=t_all,
                  sample=f"one full predict_ddim (E2D shape probe + E2D + EDT + {ddim_steps} DDIM steps + D3D), 1 sample of "
                         f"{slices}x{size}x{size}, fp32, {cores} threads; nothing scaled")
    return t_all, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    times = []
    detail = None
    for i in range(args.warmup + args.steps):  # every step, warm-up included, is one full prediction
        t, detail = cpu_reference_prediction(args.size, args.slices, args.ddim_steps)
        if i >= args.warmup:
            times.append(t)
    sec = statistics.mean(times)
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": "3D flow-field predictions/sec", "value": val, "unit": "predictions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sec, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "arm": {"samples_per_step": 1, "path": "reference op sequence on host CPU (oracle port; /root/reference is not present on the GPU box)"},
        "cpu_baseline": {"value": val, "unit": "predictions/s", "cores": detail["cores"], "kind": "port", "sample": detail["sample"]},
        "e2e": {"value": val, "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "detail": {k: v for k, v in detail.items() if k != "sample"}, "step_seconds": times, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def workload_config(args):
    """The workload both arms are quoted on (identical dict in the two JSON lines); arm-specific run details go to `arm`."""
    return {"workload": f"predict_ddim DDIM-{args.ddim_steps} end to end (E2D encode -> conditioned-UNet loop -> D3D decode) on synthetic "
                        f"microstructures of {args.slices}x{args.size}x{args.size}, UNet in17/out8 k3 zeros-pad attn 3..2 + dual-branch VAE, "
                        "random-init weights (BASELINE.json configs[2]/[3])",
            "ddim_steps": args.ddim_steps, "slices": args.slices, "size": args.size,
            "l2": "no explicit flush between timed steps: a step streams activation tensors of "
                  f"{min(args.batch_per_gpu, args.vae_chunk) * args.slices * args.size ** 2 * 128 * 2 / 1e9:.2f} GB each (GPU arm, "
                  "128-channel VAE maps) through the 126 MB L2"}


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
        sec, det = cpu_reference_prediction(args.size, args.slices, args.ddim_steps)
        cpu_base = {"value": 1.0 / sec, "unit": "predictions/s", "cores": det["cores"], "kind": "port", "sample": det["sample"]}

    S, H = args.slices, args.size
    usd, vsd = synth.synth_unet_state(seed=0), synth.synth_vae_state(seed=1)
    pred = B200LatentDiffusionPredictor("UNet", dict(synth.UNET_KWARGS), True, unet_state=usd, vae_state=vsd,
                                        norm_factors=synth.NORM_FACTORS, num_slices=S, num_timesteps=1000, precision=args.precision,
                                        use_graph=not args.no_graph, vae_chunk=args.vae_chunk, device=dev)
    del usd, vsd

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        sync_all()
        if sampler:
            sampler.mark()
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

    def measure(global_batch, per_rank, warmup, steps, sampler=None):
        """Whole-job throughput of `global_batch` samples per step: this rank's shard is `per_rank` samples (the rank-local
        slice of the synthetic global batch, no communication).  Returns resident and end-to-end numbers."""
        B = per_rank
        lo = sum(sharding.shard_range(global_batch, r, world)[1] - sharding.shard_range(global_batch, r, world)[0] for r in range(rank))
        img, v2d = synth.synth_inputs(B, num_slices=S, size=H, seed=2024 + lo)
        noise = synth.synth_noise(B, num_slices=S, latent_size=H // 4, seed=42 + lo)
        img_h, v2d_h, noise_h = img.pin_memory(), v2d.pin_memory(), noise.pin_memory()
        img_d, v2d_d, noise_d = img.to(dev), v2d.to(dev), noise.to(dev)
        n_out = global_batch if (world > 1 and rank == 0) else B  # rank 0 reads the GATHERED fields back at N > 1
        out_h = torch.empty(n_out, S, 3, H, H, dtype=torch.float32).pin_memory()
        h2d = (img_h.numel() + v2d_h.numel() + noise_h.numel()) * 4
        d2h = out_h.numel() * 4

        def step_resident():
            out = pred.predict_ddim(img_d, v2d_d, num_steps=args.ddim_steps, eta=0.0, noise=noise_d)
            return sharding.gather_predictions(out, global_batch, dst=0) if world > 1 else out

        def step_e2e():
            a = img_h.to(dev, non_blocking=True)
            b = v2d_h.to(dev, non_blocking=True)
            c = noise_h.to(dev, non_blocking=True)
            out = pred.predict_ddim(a, b, num_steps=args.ddim_steps, eta=0.0, noise=c)
            if world > 1:
                out = sharding.gather_predictions(out, global_batch, dst=0)
            if out is not None:
                out_h.copy_(out, non_blocking=True)

        if sampler:
            sampler.start()  # before the warm-up steps; timed() marks where the timed region begins
        for _ in range(warmup):
            step_resident()
        ms, launches, clocks = timed(step_resident, steps, sampler)
        for _ in range(min(2, warmup)):
            step_e2e()
        ms_e2e, _, _ = timed(step_e2e, steps)
        t_h2d = torch.tensor([float(h2d)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_h2d)  # whole-job H2D bytes per step
        return dict(ms_per_step=ms / steps, value=global_batch / (ms / steps / 1e3), e2e_ms_per_step=ms_e2e / steps,
                    e2e_value=global_batch / (ms_e2e / steps / 1e3), launches=launches, clocks=clocks,
                    h2d_bytes=int(t_h2d.item()), d2h_bytes=int(d2h if world == 1 else global_batch * S * 3 * H * H * 4), per_rank=B)

    # ---- primary measurement --------------------------------------------------------------------------------------
    strong_primary = args.global_batch > 0
    if strong_primary:
        G = args.global_batch
        lo, hi = sharding.shard_range(G, rank, world)
        B = hi - lo
        if B < 1:
            raise SystemExit(f"--global-batch {G} leaves rank {rank} of {world} without a sample")
    else:
        B = args.batch_per_gpu
        G = B * world
    warm = max(args.warmup, 3)
    m = measure(G, B, warm, args.steps, ClockSampler(local))
    ms_step, value = m["ms_per_step"], m["value"]

    # ---- stage breakdown + rooflines (measured live with CUDA events on the launch stream, primary session) ----------
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

    def time_replayed(fn, reps):
        """ms per call of fn, `reps` calls captured into one CUDA graph and replayed: the host's launch rate does not enter
        (eager per-launch timing of 20-70 us kernels swings by 25 % between boxes with the host's speed)."""
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                sp = torch.cuda.current_stream(dev).cuda_stream
                for _ in range(reps):
                    fn(sp)
        torch.cuda.current_stream(dev).wait_stream(side)
        g.replay()
        return time_launches(g.replay, 3) / reps

    nchunk = len(ses["starts"])
    t_e2d = time_launches(lambda: [ses["e2d"]["program"].run(s, variant=i) for i in range(nchunk)], 3)
    t_d3d = time_launches(lambda: [ses["d3d"]["program"].run(s, variant=i) for i in range(nchunk)], 3)
    img_d, v2d_d = ses["img"].clone(), ses["v2d"].clone()
    noise_d = torch.randn(ses["N"], 8, ses["h"], ses["w"], device=dev)
    t_cond = time_launches(lambda: pred._conditioning(ses, img_d, v2d_d, s), 3)        # copies + E2D + EDT + bilinear
    t_dec = time_launches(lambda: pred._decode(ses, s), 3)                               # D3D + the returned copy
    # the UNet step exactly as the hot path runs it: replays of the captured timestep graph (UNet body + final_conv fused
    # with the sampler update); one loop = ddim_steps replays from step index 0
    graph = ses["graph"][1] if ses.get("graph") else None

    def one_loop():
        pred._set_latent(ses, noise_d, s)
        ses["state"][0:2].zero_()
        if graph is not None:
            for _ in range(args.ddim_steps):
                graph.replay()
        else:
            for _ in range(args.ddim_steps):
                pred._one_step(ses, 1, ses["coef"], (-30.0, 30.0), s)
    t_loop = time_launches(one_loop, 3)
    t_set = time_launches(lambda: pred._set_latent(ses, noise_d, s), 3)
    t_unet = (t_loop - t_set) / args.ddim_steps
    stage_sum = t_cond + t_loop + t_dec
    # dominant kernel: the tcgen05 implicit-GEMM conv engine (~80 % of the step's device time).  Its roofline launch is
    # the heaviest single launch of the step (a D3D 3x3x3 conv), timed alone with CUDA events on the launch stream ->
    # burst peak.
    dom, dom_name = None, ""
    for name, fn in ses["d3d"]["program"].steps:
        plan = getattr(fn, "__self__", None)
        if plan is not None and hasattr(plan, "flops") and (dom is None or plan.flops > dom.flops):
            dom, dom_name = plan, name
    t_dom = time_launches(lambda: dom.run(s), 10)
    tc_ach = dom.flops / (t_dom * 1e-3) / 1e12
    di = dom.info2()
    dd = dom.desc
    dom_label = (f"conv_v2_kernel<BN={di['block_n']},halo={di['halo']}> (D3D {dom_name}: {dd.cin[0]}->{dd.cout} "
                 f"3x3x3 @ {dd.N}x{dd.D}x{dd.H}x{dd.W}, timed alone: burst peak)")
    # DRAM bytes of that launch from the committed `ncu --set full` capture of the same shape (profiles/), if present
    traffic = None
    shape_key = f"3d {dd.N} {dd.cin[0]} {dd.cout} {dd.H}"
    for tname in ("r2_dominant_kernel_ncu.json", "r1_dominant_kernel_ncu.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("shape", "").startswith(shape_key):
                traffic = tj.get("dram_traffic_bytes")
                break
    # scheduler kernel on >= 64 samples' worth of latent (369 MB > L2) for an HBM-bound number
    n_el = ELEMS_PER_SAMPLE * 64
    xs, es, zs = (torch.randn(n_el, device=dev) for _ in range(3))
    coef = pred.scheduler._ddpm_table
    t_sched = time_launches(lambda: _lib.call("b2d_scheduler_step", 0, xs.data_ptr(), es.data_ptr(), zs.data_ptr(), xs.data_ptr(),
                                              n_el, coef.data_ptr(), None, 500, 0, 1, -30.0, 30.0, None, 0, 0, 0, None, None, 0, s), 20)
    hbm_ach = 16.0 * n_el / (t_sched * 1e-3) / 1e9
    del xs, es, zs
    # training slice: the flat Adam update over the UNet's 139.8 M parameters, 28 B / parameter
    n_par = 139810952
    bufs = [torch.zeros(n_par + 8, device=dev) for _ in range(4)]
    bufs[1].normal_()
    t_adam = time_launches(lambda: _lib.call("b2d_adam_step", bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(),
                                             n_par, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, s), 10)
    adam_ach = 28.0 * n_par / (t_adam * 1e-3) / 1e9
    del bufs
    # the UNet's tensor-core launches alone (3x3 convs, transposed convs, projections): per-launch CUDA events, eager
    conv_us, conv_flops = 0.0, 0.0
    for name, fn in ses["unet"]["program"].steps:
        plan = getattr(fn, "__self__", None)
        if plan is not None and hasattr(plan, "flops"):
            conv_us += time_replayed(plan.run, 8) * 1e3
            conv_flops += plan.flops
    scale = (S / 11.0) * (H / 256.0) ** 2
    unet_slices = ses["N"]

    # ---- secondary: fixed global batch (strong scaling), micro-batched on each GPU -----------------------------------
    strong = None
    if not strong_primary and args.strong_batch > 0 and args.strong_batch >= world:
        Gs = args.strong_batch
        lo, hi = sharding.shard_range(Gs, rank, world)
        Bs = hi - lo
        if Bs == B and Gs == G:
            sm = m  # the same configuration as the primary measurement (64 samples on 8 GPUs)
            t_unet_s = t_unet
        else:
            sm = measure(Gs, Bs, 1, max(2, min(args.steps, 3)))
            ses2 = pred._session
            g2 = ses2["graph"][1] if ses2.get("graph") else None
            nz = torch.randn(ses2["N"], 8, ses2["h"], ses2["w"], device=dev)

            def loop2():
                pred._set_latent(ses2, nz, s)
                ses2["state"][0:2].zero_()
                for _ in range(args.ddim_steps):
                    g2.replay() if g2 is not None else pred._one_step(ses2, 1, ses2["coef"], (-30.0, 30.0), s)
            t_unet_s = (time_launches(loop2, 2) - time_launches(lambda: pred._set_latent(ses2, nz, s), 2)) / args.ddim_steps
            cu2, cf2 = 0.0, 0.0
            for name, fn in ses2["unet"]["program"].steps:
                plan = getattr(fn, "__self__", None)
                if plan is not None and hasattr(plan, "flops"):
                    cu2 += time_replayed(plan.run, 4) * 1e3
                    cf2 += plan.flops
            sm["conv_tflops"] = cf2 / cu2 / 1e6 if cu2 else None
        strong = {"scaling": "strong", "global_batch": Gs, "samples_per_gpu": sm["per_rank"], "value": sm["value"], "unit": "predictions/s",
                  "ms_per_step": sm["ms_per_step"], "e2e": {"value": sm["e2e_value"], "unit": "predictions/s", "ms_per_step": sm["e2e_ms_per_step"],
                                                           "h2d_bytes_per_step": sm["h2d_bytes"], "d2h_bytes_per_step": sm["d2h_bytes"]},
                  "unet_step_ms": t_unet_s, "unet_slices_per_launch": sm["per_rank"] * S,
                  "unet_tflops": sm["per_rank"] * FLOP_UNET_STEP * scale / (t_unet_s * 1e-3) / 1e12,
                  "unet_frac_of_sustained_peak": sm["per_rank"] * FLOP_UNET_STEP * scale / (t_unet_s * 1e-3) / 1e12 / peaks["tc_sustained"],
                  "unet_conv_tflops": sm.get("conv_tflops"),
                  "unet_conv_frac_of_sustained_peak": (sm["conv_tflops"] / peaks["tc_sustained"]) if sm.get("conv_tflops") else None,
                  "vae_chunk": args.vae_chunk}

    if rank == 0:
        flop_step = B * (FLOP_E2D + FLOP_D3D + args.ddim_steps * FLOP_UNET_STEP) * scale
        unet_tf = B * FLOP_UNET_STEP * scale / (t_unet * 1e-3) / 1e12
        line = {
            "metric": "3D flow-field predictions/sec", "value": value, "unit": "predictions/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong_primary else "weak",
            "vs_baseline": None,
            "dtype": {"f16": "f16 (IEEE fp16 operands, fp32 accumulate)", "bf16": "bf16",
                      "fp32x": "bf16x3 (fp32-class hi/lo split)"}[args.precision], "data": "synthetic",
            "config": workload_config(args),
            "arm": {"samples_per_gpu": B, "global_batch": G, "parallelism": f"batch-sharded x{world}, final gather only",
                    "loop": "CUDA-graphed timestep (UNet body + final_conv fused with the sampler update)" if graph is not None else "eager",
                    "vae_chunk": args.vae_chunk,
                    "l2": "working set (>= 1.4 GB per activation tensor) far exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": m["e2e_value"], "unit": "predictions/s", "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"],
                    "ms_per_step": m["e2e_ms_per_step"]},
            "gpu_launches": m["launches"],
            "clocks": m["clocks"],
            "roofline": {"bound": "tensor", "achieved": tc_ach, "peak": peaks["tc_burst"], "unit": "TFLOP/s", "frac": tc_ach / peaks["tc_burst"],
                         "traffic": traffic, "kernel": dom_label,
                         "peak_source": peaks["src"], "flops_per_launch": dom.flops, "ms_per_launch": t_dom},
            "roofline_scheduler": {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm"],
                                   "bytes_per_launch": 16.0 * n_el, "ms_per_launch": t_sched,
                                   "note": "DDPM step with host noise, 16 B/element, 64 samples (369 MB > L2)"},
            "roofline_adam": {"bound": "hbm", "achieved": adam_ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": adam_ach / peaks["hbm"],
                              "bytes_per_launch": 28.0 * n_par, "ms_per_launch": t_adam,
                              "note": "training slice: torch.optim.Adam update of the UNet's 139.8 M fp32 parameters, 28 B/parameter"},
            "stages": {"e2d_ms": t_e2d, "unet_step_ms": t_unet, "d3d_ms": t_d3d,
                       "unet_conv_us_sum": conv_us, "unet_conv_note": "each conv launch timed as 8 graph-replayed launches", "unet_conv_tflops": conv_flops / conv_us / 1e6 if conv_us else None,
                       "unet_conv_frac_of_sustained_peak": conv_flops / conv_us / 1e6 / peaks["tc_sustained"] if conv_us else None,
                       "unet_step_note": f"replayed timestep graph, {unet_slices} slice-images per launch",
                       "conditioning_ms": t_cond, "loop_ms": t_loop, "decode_ms": t_dec, "sum_ms": stage_sum,
                       "sum_over_ms_per_step": stage_sum / ms_step,
                       "e2d_tflops": B * FLOP_E2D * scale / (t_e2d * 1e-3) / 1e12,
                       "unet_tflops": unet_tf, "unet_frac_of_sustained_peak": unet_tf / peaks["tc_sustained"],
                       "d3d_tflops": B * FLOP_D3D * scale / (t_d3d * 1e-3) / 1e12,
                       "step_tflops": flop_step / (ms_step * 1e-3) / 1e12, "frac_of_sustained_peak": flop_step / (ms_step * 1e-3) / 1e12 / peaks["tc_sustained"]},
        }
        if strong is not None:
            line["strong"] = strong
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
