"""B200-native latent-diffusion sampling path (drop-in for the PyTorch path of
Ruby-004/Diffusion_model_project): UNet eps-prediction, DDPM/DDIM scheduler step, dual-branch VAE
encode/decode and the predict / predict_ddim loops, over hand-written sm_100a CUDA (libb2d.so).

Importing the package does not load the CUDA library; the first kernel call does, and raises if it
is missing -- there is no CPU or PyTorch fallback on this path.
"""
__all__ = ["B200UNet", "B200Scheduler", "B200DualVAE", "B200LatentDiffusionPredictor", "UNetTrainer", "LatentDiffusionTrainer"]


def __getattr__(name):
    if name == "B200UNet":
        from .unet import B200UNet
        return B200UNet
    if name == "B200Scheduler":
        from .scheduler import B200Scheduler
        return B200Scheduler
    if name == "B200DualVAE":
        from .vae import B200DualVAE
        return B200DualVAE
    if name == "B200LatentDiffusionPredictor":
        from .predictor import B200LatentDiffusionPredictor
        return B200LatentDiffusionPredictor
    if name in ("UNetTrainer", "LatentDiffusionTrainer"):
        from . import train
        return getattr(train, name)
    raise AttributeError(name)
