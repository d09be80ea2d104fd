"""Drop-in for Backend/hybrid/hybrid3diffusionspeed.py: same public names, libxrd.so underneath."""
import torch
from xrd_b200 import (HybridDenoisingRouter, EnhancedNAFNet, UNetDiffusion, DiffusionDenoiser,  # noqa: F401
                      NoiseAnalyzer, FusionModule)
from xrd_b200 import models as _m
NAFBlock, LayerNorm, SimpleGate, ResidualBlock, AttentionBlock = _m.NAFBlock, _m.LayerNorm, _m.SimpleGate, _m.ResidualBlock, _m.AttentionBlock

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
