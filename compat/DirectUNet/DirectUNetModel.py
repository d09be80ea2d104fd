"""Drop-in for Backend/DirectUNet/DirectUNetModel.py (the model run.py:15 imports): same public name, libxrd.so underneath.
The training-side names of that file (VGGPerceptualLoss, HybridLoss, XRayDataset, train_denoiser) are out of scope."""
import torch
from xrd_b200 import ExpertDenoiser  # noqa: F401

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # DirectUNetModel.py:10
