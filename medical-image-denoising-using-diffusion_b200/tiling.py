"""Overlap-tiled evaluation of images larger than the networks' native field
(BASELINE.json configs[4]: 1024x1024 through 512x512 tiles with halos).

The reference has no tiling code: `/denoise` resizes everything to 512x512
(RUN:197-201) and back (RUN:143-149).  Tiling serves that caller at native
resolution.  The decomposition is defined in csrc/tiles.cu / DESIGN.md section 9;
the tile stack of a (B,1,H,W) batch is itself a (B*ny*nx,1,tile,tile) batch, so
any of the drop-in models runs on it unchanged and images never interact
(sharding by image keeps every tile of an image on one GPU).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Tuple

import torch

from . import _lib
from .models import _image_arg, _ptr, _stream_ptr


def tile_plan(height: int, width: int, tile: int = 512, halo: int = 64) -> Tuple[List[int], List[int]]:
    """Tile origins along y and x (host only; xrd_tiles_plan)."""
    lib = _lib.load()
    ny, nx = C.c_int(0), C.c_int(0)
    _lib.check(lib.xrd_tiles_plan(height, width, tile, halo, C.byref(ny), C.byref(nx), None, None, 0))
    cap = max(ny.value, nx.value)
    oy, ox = (C.c_int * cap)(), (C.c_int * cap)()
    _lib.check(lib.xrd_tiles_plan(height, width, tile, halo, C.byref(ny), C.byref(nx), oy, ox, cap))
    return list(oy[:ny.value]), list(ox[:nx.value])


def extract_tiles(img: torch.Tensor, tile: int = 512, halo: int = 64) -> torch.Tensor:
    """(B,1,H,W) -> (B*ny*nx,1,tile,tile), tiles of one image contiguous, ty-major."""
    x = _image_arg(img, "img")
    B, _, H, W = x.shape
    oy, ox = tile_plan(H, W, tile, halo)
    tiles = torch.empty((B * len(oy) * len(ox), 1, tile, tile), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().xrd_tiles_extract(_ptr(x), _ptr(tiles), B, H, W, tile, halo, _stream_ptr(x.device)))
    return tiles


def blend_tiles(tiles: torch.Tensor, batch: int, height: int, width: int, tile: int = 512, halo: int = 64) -> torch.Tensor:
    """Inverse of extract_tiles with linear cross-fades over the overlaps -> (batch,1,height,width)."""
    t = _image_arg(tiles, "tiles")
    oy, ox = tile_plan(height, width, tile, halo)
    if tuple(t.shape) != (batch * len(oy) * len(ox), 1, tile, tile):
        raise _lib.XrdError(f"tiles has shape {tuple(t.shape)}, expected {(batch * len(oy) * len(ox), 1, tile, tile)}")
    out = torch.empty((batch, 1, height, width), dtype=torch.float32, device=t.device)
    _lib.check(_lib.load().xrd_tiles_blend(_ptr(t), _ptr(out), batch, height, width, tile, halo, _stream_ptr(t.device)))
    return out


@torch.no_grad()
def denoise_tiled(model: Callable[[torch.Tensor], torch.Tensor], noisy: torch.Tensor, tile: int = 512, halo: int = 64) -> torch.Tensor:
    """Run `model` (HybridDenoisingRouter, EnhancedNAFNet, or `lambda t: wrapper.denoise(t, steps)`)
    on every overlap tile of `noisy` and cross-fade the results."""
    B, _, H, W = noisy.shape
    out_tiles = model(extract_tiles(noisy, tile, halo))
    return blend_tiles(out_tiles, B, H, W, tile, halo).to(noisy.dtype)
