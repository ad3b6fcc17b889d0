"""CPU emulation of the conv engine's GEMM view (include/b2d.h: k = kbase[s] + tap*cin[s] + c,
zero-filled halo, phase-major rows for ConvTranspose2d) used to check the host-side weight
packing / tap tables without a GPU.  Test infrastructure only."""
from __future__ import annotations

import torch


def emulate_conv(inputs, pw, cout, nphase=1, stride=1):
    """inputs: list of fp32 channels-last tensors [N,D,H,W,Cpad]; pw: engine.PackedWeight (fp32 view used).
    Returns planar fp32 (N, cout, D, OH*up, OW*up)."""
    N, D, H, W, _ = inputs[0].shape
    OH, OW = H // stride, W // stride
    wmat = pw.w.float()
    rows = nphase * cout
    acc = torch.zeros(N, D, OH, OW, rows)
    segs = [(x, pw.kbase[i]) for i, x in enumerate(inputs)]
    for x, kb in segs:
        C = x.shape[-1]
        for t, (dz, dy, dx) in enumerate(pw.taps):
            # gather the tap-shifted, stride-subsampled, zero-padded input
            g = torch.zeros(N, D, OH, OW, C)
            for z in range(D):
                zz = z + dz
                if zz < 0 or zz >= D:
                    continue
                ys = torch.arange(OH) * stride + dy
                xs = torch.arange(OW) * stride + dx
                vy = (ys >= 0) & (ys < H)
                vx = (xs >= 0) & (xs < W)
                sub = x[:, zz][:, ys.clamp(0, H - 1)][:, :, xs.clamp(0, W - 1)]
                sub = sub * (vy[:, None] & vx[None, :])[None, :, :, None]
                g[:, z] = sub
            wk = wmat[:rows, kb + t * C: kb + (t + 1) * C]  # [rows, C]
            acc += g @ wk.t()
    if pw.bias is not None:
        acc = acc + pw.bias.float().repeat(nphase)
    if nphase == 1:
        return acc.permute(0, 4, 1, 2, 3).contiguous()
    out = torch.zeros(N, cout, D, 2 * OH, 2 * OW)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        out[:, :, :, py::2, px::2] = acc[..., ph * cout:(ph + 1) * cout].permute(0, 4, 1, 2, 3)
    return out
