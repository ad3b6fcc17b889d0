"""Where the weights come from: the reference's own objects and checkpoint directories.

Host-side logic only (state-dict plumbing, JSON, file discovery) -- no kernels, CPU-testable.  Mirrors
`Predictor.from_directory` (Diffusion_model/src/predictor.py:222-250), the dual-VAE branch of
`LatentDiffusionPredictor.__init__` (predictor.py:301-613: separate encoder / decoder directories, `vae_log.json`
norm factors) and `load_weights` (predictor.py:195-220).  The single-branch / conditional VAE variants and the legacy
`layers.N` key format (predictor.py:30-117) are outside the sampling path this package replaces and are refused.
"""
from __future__ import annotations

import json
import os
import os.path as osp
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

_MODEL_FILES = ("best_model.pt", "vae.pt", "model.pt")  # predictor.py:499-503 (search order of the separate-path branch)


@dataclass
class PredictorSpec:
    """Everything B200LatentDiffusionPredictor needs, in the reference's own terms."""
    model_kwargs: dict
    unet_state: Dict[str, torch.Tensor]
    vae_state: Dict[str, torch.Tensor]          # keys encoder_2d.*, decoder_3d.* (and encoder_3d.* for encode_target)
    norm_factors: List[float]
    num_slices: int = 11
    num_timesteps: int = 1000
    distance_transform: bool = True


def unet_kwargs_from_module(model) -> dict:
    """Constructor kwargs of a reference UNet instance, read back from the attributes it keeps (unet/models.py:49-64)."""
    return dict(in_channels=model.in_channels, out_channels=model.out_channels, features=list(model.features),
                kernel_size=model.kernel_size, padding_mode=model.padding_mode, activation=model._activation,
                final_activation=model._final_activation, attention=model.attention, dropout=model.dropout,
                time_embedding_dim=model.time_embedding_dim)


def spec_from_reference(predictor) -> PredictorSpec:
    """From a constructed (and weight-loaded) reference LatentDiffusionPredictor: its modules' own state_dicts.
    The reference tensors are only read (the B200 modules repack copies)."""
    if not getattr(predictor, "vae_is_dual", False):
        raise NotImplementedError("only the dual-branch VAE (E2D / D3D) is on the sampling path this package replaces "
                                  "(predictor.py:401-568); standard / conditional VAEs are out of scope")
    dt = predictor.distance_transform
    dt = bool(dt.detach().reshape(-1)[0].item()) if torch.is_tensor(dt) else bool(dt)
    scale = predictor.normalizer["output"].scale_factors
    scale = [float(v) for v in (scale.detach().reshape(-1).tolist() if torch.is_tensor(scale) else scale)]
    return PredictorSpec(model_kwargs=unet_kwargs_from_module(predictor.model),
                         unet_state={k: v.detach() for k, v in predictor.model.state_dict().items()},
                         vae_state={k: v.detach() for k, v in predictor.vae.state_dict().items()},
                         norm_factors=scale, num_slices=int(predictor.num_slices), num_timesteps=int(predictor.num_timesteps),
                         distance_transform=dt)


def _find_model_file(folder: str) -> str:
    for name in _MODEL_FILES:
        cand = osp.join(folder, name)
        if osp.exists(cand):
            return cand
    raise FileNotFoundError(f"No model file found in {folder} (looked for {', '.join(_MODEL_FILES)})")  # predictor.py:506,517


def _branch(state: Dict[str, torch.Tensor], dual_prefix: str, plain_prefix: str, what: str, path: str) -> Dict[str, torch.Tensor]:
    """predictor.py:520-566: a checkpoint holds the branch under `encoder_2d.` (dual VAE) or `encoder.` (standard VAE)."""
    for pre in (dual_prefix, plain_prefix):
        sub = {k[len(pre):]: v for k, v in state.items() if k.startswith(pre)}
        if sub:
            if any(k.startswith("layers.") for k in sub):
                raise NotImplementedError(f"{path}: legacy 'layers.N' checkpoint format (predictor.py:30-117) is not supported; "
                                          "re-save the VAE with the named-layer modules")
            return sub
    raise ValueError(f"Cannot find {what} weights in {path}. Expected keys starting with '{dual_prefix}' or '{plain_prefix}'")


def load_dual_vae_dirs(encoder_path: str, decoder_path: str):
    """E2D from the stage-2 directory, D3D + E3D from the stage-1 directory, norm factors from the decoder's
    vae_log.json (predictor.py:363-369, 476-566).  Returns (vae_state, norm_factors or None)."""
    encoder_path, decoder_path = osp.abspath(encoder_path), osp.abspath(decoder_path)
    enc_file, dec_file = _find_model_file(encoder_path), _find_model_file(decoder_path)
    enc = torch.load(enc_file, map_location="cpu", weights_only=True)
    dec = torch.load(dec_file, map_location="cpu", weights_only=True)
    vae: Dict[str, torch.Tensor] = {}
    for k, v in _branch(enc, "encoder_2d.", "encoder.", "encoder", enc_file).items():
        vae["encoder_2d." + k] = v
    for k, v in _branch(dec, "decoder_3d.", "decoder.", "decoder", dec_file).items():
        vae["decoder_3d." + k] = v
    for k, v in _branch(dec, "encoder_3d.", "encoder.", "encoder_3d", dec_file).items():
        vae["encoder_3d." + k] = v
    norm = None
    log = osp.join(decoder_path, "vae_log.json")
    if osp.exists(log):
        with open(log) as fp:
            norm = json.load(fp).get("norm_factors")
    return vae, norm


def load_directory(folder: str, *, vae_encoder_path: Optional[str] = None, vae_decoder_path: Optional[str] = None) -> PredictorSpec:
    """`Predictor.from_directory` (predictor.py:222-250): `log.json` -> predictor kwargs, VAE directories -> frozen VAE
    weights + norm factors, then `model.pt` (predictor.py:195-220) whose `model.*` entries are the UNet and whose
    `vae.*` / `normalizer.output.*` entries, when present, are what training saved alongside it."""
    with open(osp.join(folder, "log.json")) as fp:
        params = json.load(fp)["params"]
    ptype = params["training"]["predictor_type"]
    if ptype != "latent-diffusion":
        raise ValueError(f"Unknown or unsupported predictor type: {ptype}")  # predictor.py:245
    kw = dict(params["training"]["predictor"])
    if kw.get("model_name", "UNet") != "UNet":
        raise ValueError("only the 'UNet' denoiser exists in the reference (predictor.py:136)")
    model_kwargs = dict(kw.get("model_kwargs") or {})
    model_kwargs.setdefault("time_embedding_dim", 64)  # predictor.py:317-318
    enc_dir = vae_encoder_path or kw.get("vae_encoder_path")
    dec_dir = vae_decoder_path or kw.get("vae_decoder_path")
    if enc_dir is None or dec_dir is None:
        raise ValueError("VAE path must be provided for latent diffusion: this path needs both vae_encoder_path (E2D) and "
                         "vae_decoder_path (D3D / E3D) (predictor.py:342-344, 476-480)")
    vae_state, norm = load_dual_vae_dirs(enc_dir, dec_dir)
    state = torch.load(osp.join(folder, "model.pt"), map_location="cpu", weights_only=True)
    unet_state = {k[len("model."):]: v for k, v in state.items() if k.startswith("model.")}
    if not unet_state:
        raise ValueError(f"{folder}/model.pt holds no 'model.*' entries")
    saved_vae = {k[len("vae."):]: v for k, v in state.items() if k.startswith("vae.")}
    for k, v in saved_vae.items():  # load_state_dict(strict) overwrites the directory weights with the saved copy
        if k.split(".")[0] in ("encoder_2d", "encoder_3d", "decoder_3d"):
            vae_state[k] = v
    saved_scale = state.get("normalizer.output.scale_factors")
    if saved_scale is not None:
        norm = [float(v) for v in saved_scale.reshape(-1).tolist()]
    if norm is None:
        lat = model_kwargs.get("out_channels", 4)
        norm = [1.0] * lat  # predictor.py:337-340 default (and the reference's warning at :577)
    dt = state.get("distance_transform")
    dt = bool(dt.reshape(-1)[0].item()) if dt is not None else bool(kw.get("distance_transform", True))
    return PredictorSpec(model_kwargs=model_kwargs, unet_state=unet_state, vae_state=vae_state, norm_factors=[float(v) for v in norm],
                         num_slices=int(kw.get("num_slices", 11)), num_timesteps=int(kw.get("num_timesteps", 1000)),
                         distance_transform=dt)
