"""Scheduler kernel vs the CPU oracle (bit-exact: the kernel keeps the reference's fp32 operation
order) and vs golden vectors produced by the reference itself."""
import os

import numpy as np
import pytest
import torch

from diffusion_model_project_b200 import _lib
from diffusion_model_project_b200.scheduler import B200Scheduler
from oracle.scheduler import OracleScheduler

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def sch():
    return B200Scheduler(1000, device="cuda")


def test_p_sample_and_ddim_vs_golden(sch, golden_dir):
    g = np.load(os.path.join(golden_dir, "scheduler.npz"))
    x, eps, z = (torch.from_numpy(g[k]).cuda() for k in ("x", "eps", "z"))
    for t in g["p_ts"].tolist():
        got = sch.p_sample(eps, x, t, clip_denoised=True, clip_range=(-30.0, 30.0), noise=z).cpu().numpy()
        assert np.array_equal(got, g[f"p_sample_{t}"]), t
        assert np.array_equal(sch.q_sample(x, t, eps).cpu().numpy(), g[f"q_sample_{t}"]), t
        assert np.array_equal(sch.predict_x0_from_noise(x, t, eps).cpu().numpy(), g[f"x0_{t}"]), t
    assert np.array_equal(sch.p_sample(eps, x, 999, noise=z).cpu().numpy(), g["p_sample_default_clip_999"])
    for t, tp in g["ddim_pairs"].tolist():
        assert np.array_equal(sch.ddim_sample(eps, x, t, tp, 0.0).cpu().numpy(), g[f"ddim_{t}_{tp}"])
        assert np.array_equal(sch.ddim_sample(eps, x, t, tp, 0.5, noise=z).cpu().numpy(), g[f"ddim_eta05_{t}_{tp}"])


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 8 * 64 * 64 * 11, 360448 * 4 + 2])
def test_bit_exact_vs_oracle_ragged_sizes(sch, n):
    o = OracleScheduler(1000)
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, generator=g) * 4
    e = torch.randn(n, generator=g)
    z = torch.randn(n, generator=g)
    for t in (999, 321, 0):
        got = sch.p_sample(e.cuda(), x.cuda(), t, True, (-30.0, 30.0), noise=z.cuda()).cpu()
        assert torch.equal(got, o.p_sample(e, x, t, z, True, (-30.0, 30.0)))
    got = sch.ddim_sample(e.cuda(), x.cuda(), 999, 978).cpu()
    assert torch.equal(got, o.ddim_sample(e, x, 999, 978))


def test_tensor_timesteps(sch):
    o = OracleScheduler(1000)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 8, 4, 4, generator=g)
    e = torch.randn(3, 8, 4, 4, generator=g)
    t = torch.tensor([10, 500, 999])
    got = sch.q_sample(x.cuda(), t.cuda(), e.cuda()).cpu()
    assert torch.equal(got, o.q_sample(x, t, e))
    got = sch.predict_x0_from_noise(x.cuda(), t.cuda(), e.cuda()).cpu()
    assert torch.equal(got, o.predict_x0_from_noise(x, t, e))


def test_device_step_counter_and_bf16_copy(sch):
    n_pix, C, stride = 1000, 8, 64
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n_pix, C, generator=g).cuda()
    e = torch.randn(n_pix, C, generator=g).cuda()
    coef = sch.ddim_coef_rows([999, 500, 0], 0.0).cuda()
    step = torch.zeros(2, dtype=torch.int32, device="cuda")  # [0] step index, [1] caller-owned ticket word
    buf = torch.zeros(n_pix, stride, dtype=torch.bfloat16, device="cuda")
    o = OracleScheduler(1000)
    xr = x.cpu()
    ts = [999, 500, 0]
    for i in range(3):
        _lib.call("b2d_scheduler_step", 1, x.data_ptr(), e.data_ptr(), None, x.data_ptr(), x.numel(), coef.data_ptr(),
                  step.data_ptr(), 0, 1, 1, -30.0, 30.0, buf.data_ptr(), C, stride, 0, None, step.data_ptr() + 4, 0, _lib.stream_ptr())
        xr = o.ddim_sample(e.cpu(), xr, ts[i], ts[i + 1] if i < 2 else -1)
        assert step.tolist() == [i + 1, 0]  # advanced once, ticket left at zero
        assert torch.equal(x.cpu(), xr)
        assert torch.equal(buf[:, :C].float().cpu(), xr.to(torch.bfloat16).float())
        assert buf[:, C:].abs().max().item() == 0


def test_philox_noise_statistics(sch):
    n = 1 << 20
    x = torch.zeros(n, device="cuda")
    e = torch.zeros(n, device="cuda")
    coef = torch.tensor([[1.0, 0.0, 0.0, 0.0, 1.0, 0, 0, 0], [1.0, 0.0, 0.0, 0.0, 1.0, 0, 0, 0]], device="cuda")
    outs = []
    for row in (0, 1):
        out = torch.empty(n, device="cuda")
        _lib.call("b2d_scheduler_step", 0, x.data_ptr(), e.data_ptr(), None, out.data_ptr(), n, coef.data_ptr(), None, row, 0,
                  0, 0.0, 0.0, None, 0, 0, 1234, None, None, 0, _lib.stream_ptr())
        outs.append(out)
    for z in outs:  # out = 0 + 1*z
        assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1) < 5e-3
        assert abs((z ** 4).mean().item() - 3) < 0.05
    assert abs(torch.corrcoef(torch.stack(outs))[0, 1].item()) < 5e-3  # different rows -> independent streams


def test_invalid_arguments(sch):
    x = torch.zeros(16, device="cuda")
    with pytest.raises(ValueError):
        _lib.call("b2d_scheduler_step", 2, x.data_ptr(), x.data_ptr(), None, x.data_ptr(), 16, x.data_ptr(), None, 0, 0, 0, 0.0,
                  0.0, None, 0, 0, 0, None, None, 0, _lib.stream_ptr())
