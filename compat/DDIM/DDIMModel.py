"""Drop-in for Backend/DDIM/DDIMModel.py: same public names, libxrd.so underneath."""
import torch
from xrd_b200 import UNetDiffusion  # noqa: F401
from xrd_b200 import DiffusionDenoiser as _DiffusionDenoiser
from xrd_b200 import models as _m
ResidualBlock, AttentionBlock, SinusoidalPositionEmbeddings = _m.ResidualBlock, _m.AttentionBlock, _m.SinusoidalPositionEmbeddings

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # DDIM:16 (module-level global run.py imports)


class DiffusionDenoiser(_DiffusionDenoiser):
    """This file's copy of the sampler defaults to 25 steps (DDIM:269); the hybrid file's copy to 10 (HYB:401)."""

    def denoise(self, noisy_img, inference_steps=25, **kw):
        return super().denoise(noisy_img, inference_steps, **kw)
