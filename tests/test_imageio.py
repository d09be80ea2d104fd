"""Pre/post-processing of /denoise.  The reference's arithmetic is Pillow's 8-bit bicubic resampler
(Image.resize / torchvision Resize on a PIL 'L' image, RUN:146,197-201).  CPU: the coefficient tables the library builds,
driven through a numpy restatement of Pillow's two integer passes, reproduce Pillow itself bit for bit.  GPU (marked):
the CUDA kernels against Pillow, and the whole pre -> identity -> post chain against the reference's own lines."""
import numpy as np
import pytest
import torch
from PIL import Image

import xrd_b200

SIZES = [  # (H, W) -> (h, w): up, down (antialiased), mixed, identity on one axis, the served 512 target
    ((37, 53), (512, 512)), ((512, 512), (37, 53)), ((600, 400), (512, 512)), ((512, 512), (600, 400)), ((1024, 768), (512, 512)),
    ((512, 512), (1024, 768)), ((64, 512), (512, 512)), ((512, 300), (512, 512)), ((5, 7), (3, 2)), ((2048, 1500), (512, 512)),
]


def _pil_resize(a: np.ndarray, hw) -> np.ndarray:
    return np.asarray(Image.fromarray(a, mode="L").resize((hw[1], hw[0]), Image.BICUBIC))


def _pass(a: np.ndarray, out_size: int) -> np.ndarray:
    """One axis (the last) of Pillow's 8-bit resampler, from the library's table: int32 accumulate from 1<<21, >>22, clip."""
    ks, b, k = xrd_b200.resample_table(a.shape[-1], out_size)
    b = np.asarray(b, np.int64).reshape(out_size, 2)
    k = np.asarray(k, np.int64).reshape(out_size, ks)
    out = np.empty(a.shape[:-1] + (out_size,), np.uint8)
    src = a.astype(np.int64)
    for xx in range(out_size):
        x0, n = b[xx]
        ss = (1 << 21) + (src[..., x0:x0 + n] * k[xx, :n]).sum(-1)
        out[..., xx] = np.clip(ss >> 22, 0, 255)
    return out


def _restated_resize(a: np.ndarray, hw) -> np.ndarray:
    cur = a
    if a.shape[1] != hw[1]:
        cur = _pass(cur, hw[1])                                   # horizontal first, 8-bit intermediate
    if a.shape[0] != hw[0]:
        cur = _pass(np.ascontiguousarray(cur.T), hw[0]).T
    return np.ascontiguousarray(cur)


@pytest.mark.parametrize("src,dst", SIZES[:9])
def test_tables_reproduce_pillow_on_cpu(src, dst):
    rng = np.random.default_rng(src[0] * 7919 + dst[1])
    a = rng.integers(0, 256, size=src, dtype=np.uint8)
    assert np.array_equal(_restated_resize(a, dst), _pil_resize(a, dst))
    a[:] = 255                                                     # saturation: negative lobes must clip, not wrap
    assert np.array_equal(_restated_resize(a, dst), _pil_resize(a, dst))


@pytest.mark.gpu
@pytest.mark.parametrize("src,dst", SIZES)
def test_cuda_resize_equals_pillow(src, dst):
    rng = np.random.default_rng(src[1] * 31 + dst[0])
    imgs = rng.integers(0, 256, size=(2,) + src, dtype=np.uint8)
    imgs[1, : src[0] // 2] = 255
    imgs[1, src[0] // 2:] = 0                                      # a hard edge: overshoot on both sides
    got = xrd_b200.resize_bicubic_u8(torch.from_numpy(imgs).cuda(), dst).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], _pil_resize(imgs[i], dst)), i


@pytest.mark.gpu
def test_pre_and_post_equal_the_reference_lines():
    from torchvision import transforms
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, size=(700, 450), dtype=np.uint8)
    pil = Image.fromarray(a, mode="L")
    ref_in = transforms.Compose([transforms.Resize((512, 512), interpolation=transforms.InterpolationMode.BICUBIC),
                                 transforms.ToTensor()])(pil).unsqueeze(0)                       # RUN:197-201
    x = xrd_b200.preprocess_u8(torch.from_numpy(a).cuda())
    assert x.shape == (1, 1, 512, 512) and torch.equal(x.cpu(), ref_in)
    out = torch.from_numpy(rng.normal(0.5, 0.4, size=(1, 1, 512, 512)).astype(np.float32))      # values outside [0,1] included
    o = torch.clamp(out, 0, 1)                                                                    # RUN:110
    ref_img = Image.fromarray((o.squeeze(0).squeeze(0).numpy() * 255).astype("uint8"), mode="L").resize(pil.size, Image.BICUBIC)   # RUN:144-146
    got = xrd_b200.postprocess_u8(out.cuda(), pil.size).cpu().numpy()
    assert got.shape == (1, 700, 450) and np.array_equal(got[0], np.asarray(ref_img))
