"""Seeded synthetic workload shared by bench.py, the tests, the tools and the golden generator: synthetic X-ray-like inputs
 and the override of the parameters that are identities at default init.  No reference arithmetic
lives here (that is oracle/xrd_oracle.py, test infrastructure); this module only produces data."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

SD = Dict[str, Tensor]


def is_norm_param(key: str) -> bool:
    """True for the affine parameters of LayerNorm / GroupNorm modules in the
    reference's state_dict naming."""
    parts = key.split(".")
    if len(parts) < 2 or parts[-1] not in ("weight", "bias"):
        return False
    owner = parts[-2]
    if owner in ("norm1", "norm2", "norm"):                      # NAFBlock LayerNorms, attention GroupNorm
        return True
    if len(parts) >= 3:
        grand = parts[-3]
        if owner == "0" and grand in ("block1", "block2", "out_conv"):   # ResidualBlock / UNet out_conv GroupNorm
            return True
        if owner == "1" and grand in ("enc1", "enc2", "enc3", "mid", "dec3", "dec2", "conv1", "conv2"):
            return True                                              # router / fusion GroupNorms
    return False


def randomize_identity_params(sd: SD, seed: int = 99) -> None:
    """In place.  At default init every NAFBlock is the identity (beta = gamma
    = 0, HYB:149-150) and every norm's affine is (1, 0); a broken kernel would
    pass.  Overwrite them with seeded values."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        leaf = k.rsplit(".", 1)[-1]
        if leaf in ("beta", "gamma"):
            v.copy_(0.5 * torch.randn(v.shape, generator=g))
        elif is_norm_param(k):
            if leaf == "weight":
                v.copy_(1.0 + 0.2 * torch.randn(v.shape, generator=g))
            else:
                v.copy_(0.1 * torch.randn(v.shape, generator=g))


def randomize_batchnorm_params(sd: SD, seed: int = 99) -> None:
    """In place.  A fresh BatchNorm2d is (almost) the identity in eval mode (weight 1, bias 0, running_mean 0, running_var 1):
    a kernel that dropped the BatchNorm folding would pass.  Overwrite affine parameters and running statistics of every
    BatchNorm in `sd` (recognised by its ``running_mean`` sibling) with seeded values."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        if not k.endswith(".running_mean"):
            continue
        base = k[: -len("running_mean")]
        n = sd[k].shape
        sd[base + "weight"].copy_(1.0 + 0.2 * torch.randn(n, generator=g))
        sd[base + "bias"].copy_(0.1 * torch.randn(n, generator=g))
        sd[k].copy_(0.1 * torch.randn(n, generator=g))
        sd[base + "running_var"].copy_(0.5 + torch.rand(n, generator=g))


def synthetic_xray(batch: int, height: int, width: int, seed: int = 7) -> Tuple[Tensor, Tensor]:
    """(clean, noisy) synthetic grayscale X-ray-like fields in [0,1], float32
    (B,1,H,W).  clean = low-pass filtered uniform noise rescaled to [0.2,0.8];
    noisy = clamp(clean * (1 + 0.2 randn), 0, 1)  (speckle)."""
    g = torch.Generator().manual_seed(seed)
    lo_h, lo_w = max(2, height // 16), max(2, width // 16)
    base = torch.rand(batch, 1, lo_h, lo_w, generator=g)
    clean = F.interpolate(base, size=(height, width), mode="bicubic", align_corners=False)
    mn = clean.amin(dim=(2, 3), keepdim=True)
    mx = clean.amax(dim=(2, 3), keepdim=True)
    clean = 0.2 + 0.6 * (clean - mn) / (mx - mn + 1e-12)
    noisy = torch.clamp(clean * (1 + 0.2 * torch.randn(clean.shape, generator=g)), 0, 1)
    return clean.contiguous(), noisy.contiguous()
