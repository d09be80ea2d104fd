"""Drop-in for Backend/NafNet/NafnetModel.py: same public names, libxrd.so underneath."""
import torch
from xrd_b200 import EnhancedNAFNet  # noqa: F401
from xrd_b200 import models as _m
NAFBlock, LayerNorm, SimpleGate = _m.NAFBlock, _m.LayerNorm, _m.SimpleGate

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
