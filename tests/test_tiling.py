"""Overlap tiling (BASELINE configs[4]).  CPU part: the definition in oracle/xrd_oracle.py and the host-only plan
entry point of the C ABI.  GPU part (marked): extract/blend kernels bit-exact against the definition, the tiled hybrid
against oracle-per-tile + oracle blend, and the full 1024x1024 / 512-tile size through size-independent properties."""
import ctypes

import numpy as np
import pytest
import torch

import xrd_b200
from oracle import xrd_oracle as O

CASES = [  # (H, W, tile, halo)
    (1024, 1024, 512, 64), (512, 512, 512, 64), (96, 160, 64, 8), (100, 68, 64, 0), (72, 200, 64, 16), (64, 65, 64, 31),
]


def _plan(h, w, t, halo):
    lib = xrd_b200.load_library()
    ny, nx = ctypes.c_int(), ctypes.c_int()
    oy, ox = (ctypes.c_int * 64)(), (ctypes.c_int * 64)()
    rc = lib.xrd_tiles_plan(h, w, t, halo, ctypes.byref(ny), ctypes.byref(nx), oy, ox, 64)
    assert rc == 0, lib.xrd_last_error()
    return list(oy[:ny.value]), list(ox[:nx.value])


@pytest.mark.parametrize("h,w,t,halo", CASES)
def test_plan_matches_definition_and_covers(h, w, t, halo):
    oy, ox = _plan(h, w, t, halo)
    assert oy == O.tile_origins(h, t, halo) and ox == O.tile_origins(w, t, halo)
    for o, length in ((oy, h), (ox, w)):
        assert o[0] == 0 and o[-1] == length - t and all(b > a for a, b in zip(o, o[1:]))
        assert all(b - a <= t - 2 * halo for a, b in zip(o, o[1:]))          # neighbours overlap by at least 2*halo
    assert (len(oy), len(ox)) == ((3, 3) if h == 1024 else (len(oy), len(ox)))


def test_plan_rejects_bad_arguments():
    lib = xrd_b200.load_library()
    n = ctypes.c_int()
    for args in ((256, 512, 512, 64), (512, 512, 512, 256), (512, 512, 510, 0), (512, 512, 512, -1)):
        assert lib.xrd_tiles_plan(*args, ctypes.byref(n), ctypes.byref(n), None, None, 0) != 0
        assert lib.xrd_last_error()


@pytest.mark.parametrize("h,w,t,halo", CASES[2:])
def test_definition_is_a_partition_of_unity(h, w, t, halo):
    g = torch.Generator().manual_seed(h * 1000 + w)
    x = torch.rand(2, 1, h, w, generator=g)
    tiles = O.extract_tiles(x, t, halo)
    y = O.blend_tiles(tiles, 2, h, w, t, halo)
    assert (y - x).abs().max() < 3e-7                      # identical tiles cross-fade to the image itself
    wy = [O.tile_weights(k, len(O.tile_origins(h, t, halo)), t, halo) for k in range(len(O.tile_origins(h, t, halo)))]
    assert all(wk.min() > 0 and wk.max() == 1 for wk in wy)
    assert wy[0][0] == 1 and wy[-1][-1] == 1               # image borders keep full weight


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("h,w,t,halo", CASES)
def test_extract_and_blend_bit_exact(h, w, t, halo):
    b = 1 if h == 1024 else 3
    g = torch.Generator().manual_seed(h + w)
    x = torch.rand(b, 1, h, w, generator=g)
    tiles = xrd_b200.extract_tiles(x.cuda(), t, halo)
    ref_tiles = O.extract_tiles(x, t, halo)
    assert torch.equal(tiles.cpu(), ref_tiles)
    noisy_tiles = ref_tiles + 0.1 * torch.randn(ref_tiles.shape, generator=g)     # tiles that disagree in the overlaps
    got = xrd_b200.blend_tiles(noisy_tiles.cuda(), b, h, w, t, halo).cpu()
    ref = O.blend_tiles(noisy_tiles, b, h, w, t, halo)
    assert np.array_equal(got.numpy().view(np.uint32), ref.numpy().view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_tiled_hybrid_vs_oracle(mode, tol):
    import gpu_checks as G
    m, sd = G._hybrid(mode)
    m.inference_diffusion_steps = 5
    _, noisy = O.synthetic_xray(1, 96, 128, seed=11)
    t, halo = 64, 8
    ref_tiles = O.hybrid_forward(sd, O.extract_tiles(noisy, t, halo), 5, 50)
    ref = O.blend_tiles(ref_tiles, 1, 96, 128, t, halo)
    got = xrd_b200.denoise_tiled(m, noisy.to(G.DEV), t, halo).cpu()
    assert (got - ref).abs().max() < tol


@pytest.mark.gpu
def test_config5_size_properties():
    """1024x1024 through 512 tiles with 64-pixel halos, sampler built with noise_steps=100 (BASELINE configs[4]); the CPU
    oracle needs minutes per tile at this size, so: where one tile alone covers a pixel the blend must return that tile's
    value bit for bit, everywhere the result must lie between the covering tiles' values, and tiles of a batch must not interact."""
    import gpu_checks as G
    m, sd = G._hybrid("fp16")
    m.diffusion_wrapper = xrd_b200.DiffusionDenoiser(m.diffusion_unet, noise_steps=100)
    m.inference_diffusion_steps = 20                               # 20 of the 100 timesteps keep the test short
    _, noisy = O.synthetic_xray(1, 1024, 1024, seed=13)
    x = noisy.to(G.DEV)
    tiles_in = xrd_b200.extract_tiles(x, 512, 64)
    assert tiles_in.shape == (9, 1, 512, 512)
    tiles_out = m(tiles_in)
    y = xrd_b200.blend_tiles(tiles_out, 1, 1024, 1024, 512, 64)
    assert torch.isfinite(y).all()
    assert torch.equal(y[0, 0, :384, :384], tiles_out[0, 0, :384, :384])          # only tile (0,0) covers this block
    assert torch.equal(y[0, 0, 896:, 896:], tiles_out[8, 0, 384:, 384:])          # only tile (2,2)
    lo = torch.minimum(tiles_out[0, 0, :384, 384:], tiles_out[1, 0, :384, :128])
    hi = torch.maximum(tiles_out[0, 0, :384, 384:], tiles_out[1, 0, :384, :128])
    band = y[0, 0, :384, 384:512]
    assert (band >= lo - 1e-6).all() and (band <= hi + 1e-6).all()
    alone = m(tiles_in[4:5])                                                      # centre tile alone == inside the batch
    assert (alone - tiles_out[4:5]).abs().max() < 5e-3
    assert (xrd_b200.denoise_tiled(m, x, 512, 64) - y).abs().max() < 5e-3
