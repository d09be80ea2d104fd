"""Pins oracle/xrd_oracle.py (the CPU restatement) against vectors produced by running the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import xrd_oracle as O
from conftest import seeded_state_dict

TOL = 2e-5   # fp32 vs fp32, different op order / threading (reference itself is not bit-stable across threads)


@pytest.fixture(scope="module")
def hyb():
    return seeded_state_dict("hybrid")[1]


def test_schedule_known_answers(meta):
    for key, want in meta["schedule"].items():
        ns, st = map(int, key.split(","))
        assert O.ddim_timesteps(ns, st) == want
    assert len(O.ddim_timesteps(50, 8)) == 9 and len(O.ddim_timesteps(50, 100)) == 50


def test_tables_match_reference(meta):
    _, a, ah = O.ddim_tables(50)
    assert torch.allclose(a, torch.tensor(meta["alpha"]), rtol=0, atol=0)
    assert torch.allclose(ah, torch.tensor(meta["alpha_hat"]), rtol=0, atol=1e-7)


def test_nafnet_config1_256(golden):
    g = golden("nafnet_256_b1.npz")
    _, sd = seeded_state_dict("nafnet")
    out = O.nafnet_forward(sd, g["noisy"])
    assert (out - g["out"]).abs().max() < TOL


def test_nafnet_ragged_padding(golden):
    g = golden("nafnet_40x56_b2.npz")
    _, sd = seeded_state_dict("nafnet")
    out = O.nafnet_forward(sd, g["noisy"])
    assert out.shape == g["out"].shape
    assert (out - g["out"]).abs().max() < TOL


def test_ddim_standalone_trace(golden):
    g = golden("ddim_32_b1_s8.npz")
    _, sd = seeded_state_dict("unet")
    tr = {}
    out = O.ddim_denoise(sd, g["noisy"], 8, 50, trace=tr)
    assert len(tr["eps"]) == 9
    for n in range(9):
        assert (tr["eps"][n] - g["eps"][n]).abs().max() < 5e-5
        assert (tr["x_in"][n] - g["x_in"][n]).abs().max() < 5e-5
    assert (out - g["out"]).abs().max() < 5e-5


def test_unet_teacher_forced_eps(golden, hyb):
    g = golden("hybrid_64_b2_s50.npz")
    ts = O.ddim_timesteps(50, 50)
    for j, n in enumerate(g["keep"].tolist()):
        t = torch.full((2,), ts[n], dtype=torch.long)
        eps = O.unet_forward(hyb, g["x_in"][j], g["noisy"], t, prefix="diffusion_unet.")
        assert (eps - g["eps"][j]).abs().max() < 5e-5, n


def test_router_fusion(golden, hyb):
    g = golden("hybrid_64_b2_s50.npz")
    mask = O._sanitize(O.router_forward(hyb, g["noisy"], "router."))
    assert (mask - g["mask"]).abs().max() < 1e-6
    fused = O.fusion_forward(hyb, g["naf"], g["diff"], g["mask"], "fusion.")
    assert (fused - g["fused"]).abs().max() < TOL


def test_hybrid_end_to_end_free_running(golden, hyb):
    g = golden("hybrid_64_b2_s50.npz")
    parts = {}
    fused = O.hybrid_forward(hyb, g["noisy"], 50, 50, parts=parts)
    assert (parts["naf"] - g["naf"]).abs().max() < TOL
    assert (parts["diff"] - g["diff"]).abs().max() < 2e-4      # 50 free-running steps
    assert (fused - g["fused"]).abs().max() < 2e-4


def test_fp64_oracle_agrees(golden, hyb):
    """The fp64 variant (used to arbitrate fp32-vs-CUDA disagreements) stays within fp32 noise."""
    g = golden("hybrid_64_b2_s50.npz")
    t = torch.full((2,), 49, dtype=torch.long)
    eps = O.unet_forward(hyb, g["x_in"][0].double(), g["noisy"].double(), t, prefix="diffusion_unet.")
    assert (eps.float() - g["eps"][0]).abs().max() < 5e-5
