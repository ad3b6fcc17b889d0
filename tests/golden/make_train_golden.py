"""Golden vectors for ONE UNet training step, produced by the UNMODIFIED reference modules
imported from /root/reference: UNet (src/unet/models.py), DiffusionScheduler.q_sample (src/diffusion.py), the default
criterion normalized_mse_loss_per_component (src/unet/metrics.py) and torch.optim.Adam as configured by train.py:144-148.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_train_golden.py

Output (committed): tests/golden/train_step.npz -- loss, noise prediction, the L2 norm of every parameter gradient, a few
full gradient tensors and their post-Adam parameter deltas.  Weights and inputs are regenerated from synth on both sides.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.environ.get("B2D_GOLDEN_OUT", HERE)  # where the vectors are written (tests/test_golden_regen.py redirects it)
sys.path.insert(0, ROOT)
sys.path[:0] = ["/root/reference", "/root/reference/Diffusion_model"]
os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True

from diffusion_model_project_b200 import synth  # noqa: E402

FULL = ("final_conv.weight", "final_conv.bias", "encoder.0.0.block1.conv.weight", "encoder.0.0.time_mlp.1.weight",
        "time_mlp.0.weight", "bottleneck.block2.norm.weight", "encoder.2.1.mha.in_proj_weight", "decoder.4.0.conv.weight",
        "decoder.0.2.proj_out.bias")


def train_inputs():
    """Shared with tests/test_oracle_golden.py: N = 2 slices of a 32x32 latent (the smallest size five max-pools allow)."""
    g = torch.Generator().manual_seed(31)
    x_start = torch.randn(2, 8, 32, 32, generator=g)
    cond = torch.randn(2, 8, 32, 32, generator=g)
    feats = torch.rand(2, 1, 32, 32, generator=g)
    noise = torch.randn(2, 8, 32, 32, generator=g)
    t = torch.tensor([741, 12], dtype=torch.long)
    return x_start, cond, feats, noise, t


def main():
    from src.unet.models import UNet
    from src.diffusion import DiffusionScheduler
    from src.unet.metrics import cost_function

    torch.manual_seed(0)
    unet = UNet(**synth.UNET_KWARGS)
    unet.load_state_dict(synth.synth_unet_state(seed=0))
    unet.train()  # dropout p = 0: identical to eval, but this is what helper.py:273 does
    sch = DiffusionScheduler(1000, device="cpu")
    criterion = cost_function("normalized_mse_loss_per_component")
    opt = torch.optim.Adam(unet.parameters(), lr=1e-4, weight_decay=0.0)
    x_start, cond, feats, noise, t = train_inputs()
    before = {k: v.detach().clone() for k, v in unet.named_parameters()}
    x_t = sch.q_sample(x_start, t, noise)
    pred = unet(torch.cat([x_t, cond, feats], dim=1), t)
    loss = criterion(output=pred, target=noise)
    opt.zero_grad()
    loss.backward()
    opt.step()
    out = {"loss": np.float64(loss.item()), "pred": pred.detach().numpy(), "t": t.numpy()}
    names, norms = [], []
    for k, p in unet.named_parameters():
        names.append(k)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    params = dict(unet.named_parameters())
    for k in FULL:
        out[f"grad::{k}"] = params[k].grad.numpy()
        out[f"delta::{k}"] = (params[k].detach() - before[k]).numpy()
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), **out)
    print("loss", loss.item(), "params", len(names), "max grad norm", max(norms))


def target_inputs():
    """Shared with the tests: a 3D velocity target of 2 samples x 3 slices at 32x32, physical units."""
    g = torch.Generator().manual_seed(37)
    return torch.randn(2, 3, 3, 32, 32, generator=g) * torch.tensor(synth.NORM_FACTORS).view(1, 1, 3, 1, 1)


def main_encode_target():
    """LatentDiffusionPredictor.encode_target (predictor.py:1042-1085) of the unmodified
    reference with the seeded E3D weights -> tests/golden/encode_target.npz."""
    import tempfile
    sys.path.insert(0, HERE)
    from make_golden import build_reference_predictor
    vsd = synth.synth_vae_state(seed=1, branches=("encoder_2d", "encoder_3d", "decoder_3d"))
    with tempfile.TemporaryDirectory() as tmp:
        pred = build_reference_predictor(tmp, num_slices=3)
    pred.vae.encoder_3d.load_state_dict({k[len("encoder_3d."):]: v for k, v in vsd.items() if k.startswith("encoder_3d.")})
    with torch.no_grad():
        lat = pred.encode_target(target_inputs())
    np.savez_compressed(os.path.join(OUT, "encode_target.npz"), latents=lat.numpy())
    print("encode_target", tuple(lat.shape), lat.abs().max().item())


def field_inputs():
    """Shared with the tests: ONE sample of 2 slices at 128x128 (latent 32x32, the smallest five max-pools allow): mask,
    2D velocity (synth), a 3D velocity target in physical units, the injected noise, and the seed under which the
    reference's forward() draws its timesteps."""
    img, v2d = synth.synth_inputs(1, num_slices=2, size=128, seed=2024)
    g = torch.Generator().manual_seed(53)
    target = torch.randn(1, 2, 3, 128, 128, generator=g) * torch.tensor(synth.NORM_FACTORS).view(1, 1, 3, 1, 1) * img
    noise = torch.randn(1, 2, 8, 32, 32, generator=g)
    return img, v2d, target, noise, 1234


def main_train_from_fields():
    """The 'latent-diffusion' branch of the reference's training loop body (helper.py:277-430 with its defaults: no physics /
    velocity loss) on the UNMODIFIED reference predictor: target latents = encode_target(targets) (frozen E3D),
    preds, noise = predictor(img, velocity_2d, x_start=latents, noise=noise)  (predictor.py:636-751: frozen E2D conditioning,
    EDT features, randint timesteps, q_sample, UNet), criterion, backward, Adam -> tests/golden/train_from_fields.npz."""
    import tempfile
    sys.path.insert(0, HERE)
    from make_golden import build_reference_predictor
    from src.unet.metrics import cost_function
    vsd = synth.synth_vae_state(seed=1, branches=("encoder_2d", "encoder_3d", "decoder_3d"))
    with tempfile.TemporaryDirectory() as tmp:
        pred = build_reference_predictor(tmp, num_slices=2)
    pred.vae.encoder_3d.load_state_dict({k[len("encoder_3d."):]: v for k, v in vsd.items() if k.startswith("encoder_3d.")})
    img, v2d, target, noise, seed = field_inputs()
    pred.train()                                   # helper.py:273
    for p in pred.vae.parameters():                # the VAE is frozen (predictor.py:568-571)
        assert not p.requires_grad
    criterion = cost_function("normalized_mse_loss_per_component")
    opt = torch.optim.Adam([p for p in pred.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.0)
    before = {k: v.detach().clone() for k, v in pred.model.named_parameters()}
    with torch.enable_grad():
        latents = pred.encode_target(target, v2d)  # helper.py:288
        torch.manual_seed(seed)                    # forward() draws t = randint(...) from the default generator (:736)
        preds, target_noise = pred(img, v2d, x_start=latents, noise=noise)
        loss = criterion(output=preds, target=target_noise)
        opt.zero_grad()
        loss.backward()
        opt.step()
    torch.manual_seed(seed)
    t = torch.randint(0, pred.num_timesteps, (2,)).long()
    out = {"loss": np.float64(loss.item()), "pred": preds.detach().numpy(), "t": t.numpy(), "latents": latents.detach().numpy()}
    names, norms = [], []
    for k, p in pred.model.named_parameters():
        names.append(k)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    params = dict(pred.model.named_parameters())
    for k in ("final_conv.weight", "encoder.0.0.block1.conv.weight", "bottleneck.block2.norm.weight"):
        out[f"grad::{k}"] = params[k].grad.numpy()
        out[f"delta::{k}"] = (params[k].detach() - before[k]).numpy()
    np.savez_compressed(os.path.join(OUT, "train_from_fields.npz"), **out)
    print("train_from_fields: loss", loss.item(), "t", t.tolist(), "max grad norm", max(norms))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "step"):
        main()
    if which in ("all", "target"):
        main_encode_target()
    if which in ("all", "fields"):
        main_train_from_fields()
