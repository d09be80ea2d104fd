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


# ------------------------------------------------------------------ ExpertDenoiser
def test_expert_oracle_matches_reference(golden):
    g = golden("expert_64_b2.npz")
    _, sd = seeded_state_dict("expert")
    for x, want in ((g["noisy"], g["out"]), (g["noisy_r"], g["out_r"])):
        out = O.expert_forward(sd, x)
        assert out.shape == want.shape
        assert (out - want).abs().max() < TOL
    _, sd32 = seeded_state_dict("expert32")
    assert (O.expert_forward(sd32, g["noisy_r"]) - g["out_r_base32"]).abs().max() < TOL


def test_expert_state_dict_keys_and_seeded_init_equal_the_reference():
    import json
    import os
    import xrd_b200
    from conftest import GOLDEN
    meta = json.load(open(os.path.join(GOLDEN, "meta_expert.json")))
    torch.manual_seed(1234)
    m = xrd_b200.ExpertDenoiser(in_channels=1, base_channels=64)
    sd = m.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == meta["expert_keys"]      # incl. running stats and num_batches_tracked
    for k, (shape, s1, s2) in meta["expert_init_digest"].items():
        d = sd[k].double()
        assert list(sd[k].shape) == shape
        assert abs(float(d.sum()) - s1) <= 1e-9 * max(1.0, abs(s1)) and abs(float((d * d).sum()) - s2) <= 1e-9 * max(1.0, abs(s2)), k
