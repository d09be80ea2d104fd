"""Golden vectors for ExpertDenoiser, produced by RUNNING THE UNMODIFIED REFERENCE class
(/root/reference/Backend/DirectUNet/DirectUNetModel.py:160-255).  Build container only.  Kept apart from make_golden.py so
that the round-1 vectors need not be regenerated (that script rewrites every file).

    python tests/golden/make_golden_expert.py

Seeds: weights 1234 (after import: the module re-seeds to 42 at import, DirectUNetModel.py:11), BatchNorm affine + running
statistics override 99 (a fresh BatchNorm is the identity in eval mode), inputs 7 / 8.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import xrd_oracle as O  # noqa: E402  (seed helpers + synthetic inputs only)
from make_golden import digest, import_reference  # noqa: E402


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    import_reference()                      # installs the skimage / matplotlib stand-ins and the sys.path entry
    import DirectUNet.DirectUNetModel as EXP
    torch.manual_seed(1234)
    m = EXP.ExpertDenoiser(in_channels=1, base_channels=64).eval()
    sd = m.state_dict()
    meta = {"expert_init_digest": digest({k: v for k, v in sd.items() if v.dtype.is_floating_point}),
            "expert_keys": {k: list(v.shape) for k, v in sd.items()}}
    O.randomize_batchnorm_params(sd, 99)    # in place -> the module's parameters and buffers
    _, x1 = O.synthetic_xray(2, 64, 64, seed=7)
    _, x2 = O.synthetic_xray(1, 40, 56, seed=8)         # H, W multiples of 4 only
    y1, y2 = m(x1), m(x2)
    # a smaller base width exercises other channel counts of the same code
    torch.manual_seed(1234)
    m32 = EXP.ExpertDenoiser(in_channels=1, base_channels=32).eval()
    O.randomize_batchnorm_params(m32.state_dict(), 99)
    y3 = m32(x2)
    np.savez_compressed(os.path.join(HERE, "expert_64_b2.npz"), noisy=x1.numpy(), out=y1.numpy(), noisy_r=x2.numpy(), out_r=y2.numpy(),
                        out_r_base32=y3.numpy())
    with open(os.path.join(HERE, "meta_expert.json"), "w") as f:
        json.dump(meta, f)
    print("expert_64_b2.npz", os.path.getsize(os.path.join(HERE, "expert_64_b2.npz")), "bytes; |out| max", float(y1.abs().max()))


if __name__ == "__main__":
    main()
