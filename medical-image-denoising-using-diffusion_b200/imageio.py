"""The pixel work either side of the networks in `/denoise`, on the GPU.

    x   = preprocess_u8(img_u8)                 # RUN:191-201: Resize((512,512), BICUBIC) on the 'L' image + ToTensor()
    out = model(x)
    png_ready = postprocess_u8(out, (W0, H0))   # RUN:110,143-146: clamp, (v*255).astype('uint8'), resize(original_size, BICUBIC)

Bit-exact with the reference's CPU path, whose arithmetic is Pillow's 8-bit resampler (csrc/resample.cu).  PNG and base64
encoding stay on the host (RUN:147-149)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch

from . import _lib
from .models import _ptr, _stream_ptr


def _u8_images(img: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(img, torch.Tensor) or img.dtype != torch.uint8 or img.dim() not in (2, 3):
        raise _lib.XrdError(f"{name} must be a (H,W) or (N,H,W) uint8 tensor")
    if img.device.type != "cuda":
        raise _lib.XrdError(f"{name} is on {img.device}: this path runs on CUDA only (no CPU fallback)")
    x = img.contiguous()
    return x.unsqueeze(0) if x.dim() == 2 else x


def resize_bicubic_u8(img: torch.Tensor, size_hw: Tuple[int, int]) -> torch.Tensor:
    """PIL `Image.resize((w, h), Image.BICUBIC)` for mode 'L' images: (N,H,W) uint8 -> (N,h,w) uint8."""
    x = _u8_images(img, "img")
    n, h, w = x.shape
    ho, wo = int(size_hw[0]), int(size_hw[1])
    out = torch.empty((n, ho, wo), dtype=torch.uint8, device=x.device)
    tmp = torch.empty((n, h, wo), dtype=torch.uint8, device=x.device) if (h != ho and w != wo) else None
    _lib.check(_lib.load().xrd_resize_bicubic_u8(_ptr(x), _ptr(out), _ptr(tmp), n, h, w, ho, wo, _stream_ptr(x.device)))
    return out


def preprocess_u8(img: torch.Tensor, size: int = 512) -> torch.Tensor:
    """(H,W) or (N,H,W) uint8 grayscale -> (N,1,size,size) float32 in [0,1] (RUN:197-201)."""
    r = resize_bicubic_u8(img, (size, size))
    out = torch.empty(r.shape, dtype=torch.float32, device=r.device)
    _lib.check(_lib.load().xrd_u8_to_unit(_ptr(r), _ptr(out), C.c_int64(r.numel()), _stream_ptr(r.device)))
    return out.unsqueeze(1)


def postprocess_u8(out: torch.Tensor, original_size_wh: Tuple[int, int]) -> torch.Tensor:
    """(N,1,H,W) float model output -> (N,h0,w0) uint8; `original_size_wh` is PIL's (width, height) as run.py passes it."""
    if not isinstance(out, torch.Tensor) or out.dim() != 4 or out.shape[1] != 1 or out.device.type != "cuda":
        raise _lib.XrdError("out must be a (N,1,H,W) CUDA tensor")
    x = out.detach().to(torch.float32).contiguous()
    u8 = torch.empty((x.shape[0], x.shape[2], x.shape[3]), dtype=torch.uint8, device=x.device)
    _lib.check(_lib.load().xrd_unit_to_u8(_ptr(x), _ptr(u8), C.c_int64(x.numel()), _stream_ptr(x.device)))
    w0, h0 = int(original_size_wh[0]), int(original_size_wh[1])
    return resize_bicubic_u8(u8, (h0, w0))


def resample_table(in_size: int, out_size: int):
    """Host only: (ksize, bounds[out][2], kk[out][ksize]) of one axis, as Pillow builds it."""
    lib = _lib.load()
    ks = C.c_int(0)
    _lib.check(lib.xrd_resample_table(in_size, out_size, C.byref(ks), None, None, 0))
    b = (C.c_int * (2 * out_size))()
    k = (C.c_int * (ks.value * out_size))()
    _lib.check(lib.xrd_resample_table(in_size, out_size, C.byref(ks), b, k, ks.value * out_size))
    return ks.value, list(b), list(k)
