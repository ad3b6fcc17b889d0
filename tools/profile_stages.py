"""Per-launch CUDA-event timing of the three launch programs of one predict_ddim call (E2D, one UNet
step, D3D) at B samples of 11x256x256, eager (no graph).  usage: python tools/profile_stages.py [B] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import synth  # noqa: E402
from diffusion_model_project_b200.predictor import B200LatentDiffusionPredictor  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
pred = B200LatentDiffusionPredictor("UNet", dict(synth.UNET_KWARGS), True, unet_state=synth.synth_unet_state(seed=0),
                                    vae_state=synth.synth_vae_state(seed=1), norm_factors=synth.NORM_FACTORS, num_slices=11,
                                    num_timesteps=1000, precision="f16", device=dev)
img, v2d = synth.synth_inputs(B, num_slices=11, size=256, seed=2024)
noise = synth.synth_noise(B, num_slices=11, latent_size=64, seed=42)
pred.predict_ddim(img.to(dev), v2d.to(dev), num_steps=2, eta=0.0, noise=noise.to(dev))
ses = pred._session
s = torch.cuda.current_stream().cuda_stream
for stage in ("e2d", "unet", "d3d"):
    prog = ses[stage]["program"]
    plans = [k for k in ses[stage]["keep"] if hasattr(k, "info")]
    names = [n for n, _ in prog.steps]
    acc = [0.0] * len(names)
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        evs[0].record()
        for i, (_, fn) in enumerate(prog.steps):
            fn = fn[0] if isinstance(fn, list) else fn   # chunk variants: time the first chunk's binding
            fn(s)
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i in range(len(names)):
            acc[i] += evs[i].elapsed_time(evs[i + 1]) * 1e3 / reps
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        prog.run(s)
    b.record()
    torch.cuda.synchronize()
    total = a.elapsed_time(b) * 1e3 / reps
    print(f"== {stage} B={B}: {total:.1f} us back-to-back, {sum(acc):.1f} us summed, {len(names)} launches, {prog.flops / total / 1e6:.1f} TFLOP/s")
    pi = 0
    for n, t in zip(names, acc):
        extra = ""
        fn = dict(prog.steps)[n]
        fn = fn[0] if isinstance(fn, list) else fn
        owner = getattr(fn, "__self__", None)
        if owner is not None and hasattr(owner, "flops"):
            i2 = owner.info2(); extra = f"  {owner.flops / t / 1e6:7.1f} TFLOP/s  e{i2["engine"]} halo{i2["halo"]} ks{i2["ksplit"]} units{i2["units"]} ctas{i2["ctas"]} bn{i2["block_n"]} kg{i2["kgroups"]}"
        print(f"{t:10.1f} us  {n}{extra}")
