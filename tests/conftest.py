import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def meta():
    with open(os.path.join(GOLDEN, "meta.json")) as f:
        return json.load(f)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def seeded_state_dict(kind: str):
    """The seeded weights every parity test uses (weights seed 1234, identity-param override 99),
    built from this repo's own drop-in classes; test_host_logic proves they equal the reference's."""
    import xrd_b200
    from oracle import xrd_oracle as O
    torch.manual_seed(1234)
    if kind == "hybrid":
        m = xrd_b200.HybridDenoisingRouter({}, {}, inference_diffusion_steps=50)
    elif kind == "unet":
        m = xrd_b200.UNetDiffusion()
    elif kind == "nafnet":
        m = xrd_b200.EnhancedNAFNet()
    elif kind in ("expert", "expert32"):
        m = xrd_b200.ExpertDenoiser(in_channels=1, base_channels=64 if kind == "expert" else 32)
    else:
        raise KeyError(kind)
    m.eval()
    sd = m.state_dict()
    if kind.startswith("expert"):
        O.randomize_batchnorm_params(sd, 99)
    else:
        O.randomize_identity_params(sd, 99)
    return m, sd
