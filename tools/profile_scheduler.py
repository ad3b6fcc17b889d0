"""HBM roofline of the fused scheduler step: DDPM step with host noise on 64 samples' latents (369 MB > L2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_model_project_b200 import _lib  # noqa: E402
from diffusion_model_project_b200.scheduler import B200Scheduler  # noqa: E402

n_el = 360448 * 64
dev = "cuda"
sch = B200Scheduler(num_timesteps=1000, device=dev)
xs, es, zs = (torch.randn(n_el, device=dev) for _ in range(3))
s = torch.cuda.current_stream().cuda_stream
coef = sch._ddpm_table
for kind, nz, bpe in ((0, zs, 16.0), (1, None, 12.0)):
    tab = coef if kind == 0 else sch.ddim_coef_rows([999, 500], 0.0).to(dev)
    row = 500 if kind == 0 else 0
    fn = lambda: _lib.call("b2d_scheduler_step", kind, xs.data_ptr(), es.data_ptr(), _lib.ptr(nz), xs.data_ptr(), n_el, tab.data_ptr(), None,
                           row, 0, 1, -30.0, 30.0, None, 0, 0, 0, None, None, 0, s)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"scheduler kind={kind} {bpe:.0f} B/elem: {ms * 1e3:.1f} us, {bpe * n_el / ms / 1e6:.0f} GB/s")
