"""Drop-in for Backend/DDIM/DDIMModel.py: same public names, libxrd.so underneath."""
import torch
from xrd_b200 import UNetDiffusion, DiffusionDenoiser  # noqa: F401
from xrd_b200 import models as _m
ResidualBlock, AttentionBlock, SinusoidalPositionEmbeddings = _m.ResidualBlock, _m.AttentionBlock, _m.SinusoidalPositionEmbeddings

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # DDIM:16 (module-level global run.py imports)
