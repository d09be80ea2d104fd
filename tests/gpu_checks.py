"""GPU parity checks shared by the pytest suite (tests/test_gpu_*.py) and the diagnostic runner
(tools/gpu_diag.py).  Every function returns a dict of measured errors; the callers apply tolerances.
All calls go through the C ABI (ctypes -> libxrd.so); the oracle is only the checker."""
from __future__ import annotations

import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import xrd_b200  # noqa: E402
from xrd_b200 import _lib  # noqa: E402
from oracle import xrd_oracle as O  # noqa: E402
from conftest import load_golden, seeded_state_dict  # noqa: E402

DEV = "cuda:0"
MODES = {"bf16": _lib.MODE_BF16, "fp32": _lib.MODE_FP32_CHECK, "fp16": _lib.MODE_FP16}


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


class OpHandle:
    """A bare libxrd handle for the kernel-level hooks (no weights needed)."""

    def __init__(self, mode="bf16"):
        self.lib = _lib.load()
        cfg = _lib.default_config()
        self.h = C.c_void_p()
        _lib.check(self.lib.xrd_create(0, C.byref(cfg), C.byref(self.h)))
        _lib.check(self.lib.xrd_set_mode(self.h, MODES[mode]))

    def close(self):
        if self.h:
            self.lib.xrd_destroy(self.h)
            self.h = None

    def conv2d(self, x, w, b, k, stride, pad, impl):
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty(B, Cout, Ho, Wo, device=x.device, dtype=torch.float32)
        _lib.check(self.lib.xrd_op_conv2d(self.h, impl, _p(x), _p(w), _p(b), _p(y), B, Cin, H, W, Cout, k, stride, pad, None))
        torch.cuda.synchronize()
        return y

    def conv2d_stats(self, x, w, b, k, stride, pad, impl):
        """conv + the per-(image, 8 channel groups) sum / sum of squares its epilogue accumulates"""
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty(B, Cout, Ho, Wo, device=x.device, dtype=torch.float32)
        st = torch.zeros(B, 8, 2, device=x.device, dtype=torch.float64)
        _lib.check(self.lib.xrd_op_conv2d_stats(self.h, impl, _p(x), _p(w), _p(b), _p(y), _p(st), B, Cin, H, W, Cout, k, stride, pad, None))
        torch.cuda.synchronize()
        return y, st

    def time_last(self, iters=20):
        ms = C.c_float()
        _lib.check(self.lib.xrd_op_time_last(self.h, iters, C.byref(ms), None))
        return ms.value

    def groupnorm_act(self, x, g, b, groups, act):
        B, Cc, H, W = x.shape
        y = torch.empty_like(x)
        _lib.check(self.lib.xrd_op_groupnorm_act(self.h, _p(x), _p(g), _p(b), _p(y), B, Cc, H, W, groups, act, None))
        torch.cuda.synchronize()
        return y

    def attention(self, qkv, heads, d, impl):
        B, _, H, W = qkv.shape
        y = torch.empty(B, heads * d, H, W, device=qkv.device, dtype=torch.float32)
        _lib.check(self.lib.xrd_op_attention(self.h, impl, _p(qkv), _p(y), B, heads, d, H, W, None))
        torch.cuda.synchronize()
        return y


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


# ------------------------------------------------------------------ kernel-level
CONV_CASES_SIMT = [
    # (B, Cin, H, W, Cout, k, stride, pad)
    (2, 48, 24, 40, 48, 3, 1, 1), (1, 2, 33, 17, 48, 3, 1, 1), (2, 3, 16, 16, 48, 3, 1, 1), (1, 1, 20, 20, 32, 3, 1, 1),
    (2, 96, 16, 16, 96, 3, 2, 1), (1, 32, 16, 24, 64, 2, 2, 0), (2, 192, 8, 8, 576, 1, 1, 0), (1, 100, 9, 11, 36, 1, 1, 0),
    (1, 48, 12, 12, 1, 3, 1, 1),
]
CONV_CASES_TC = [
    (2, 48, 32, 32, 48, 3, 1, 1), (1, 96, 16, 48, 96, 3, 1, 1), (2, 144, 16, 16, 144, 3, 1, 1), (1, 192, 8, 8, 192, 3, 1, 1),
    (1, 384, 16, 16, 192, 3, 1, 1), (1, 288, 24, 24, 96, 3, 1, 1), (2, 64, 20, 28, 64, 3, 1, 1), (1, 192, 16, 16, 576, 1, 1, 0),
    (1, 512, 8, 8, 1024, 1, 1, 0), (2, 32, 16, 16, 64, 1, 1, 0), (1, 96, 32, 32, 96, 3, 2, 1), (1, 48, 64, 64, 48, 3, 2, 1),
    (1, 64, 16, 16, 128, 2, 2, 0), (1, 16, 8, 8, 32, 1, 1, 0),
]


CONV_CASES_HALO = [   # 3x3/s1/p1, W % 128 == 0 (persistent halo kernel)
    (1, 48, 8, 128, 48, 3, 1, 1), (2, 96, 6, 256, 96, 3, 1, 1), (1, 144, 5, 128, 144, 3, 1, 1), (1, 192, 4, 128, 96, 3, 1, 1),
    (1, 96, 7, 128, 48, 3, 1, 1), (3, 48, 16, 256, 96, 3, 1, 1), (1, 288, 4, 128, 144, 3, 1, 1), (1, 48, 40, 512, 48, 3, 1, 1),
]


CONV_CASES_HALO_CAT = [   # same kernel over a virtual concat of the two channel halves (B = 1)
    (1, 96, 7, 128, 48, 3, 1, 1), (1, 192, 6, 256, 96, 3, 1, 1), (1, 288, 5, 128, 144, 3, 1, 1), (1, 192, 9, 128, 48, 3, 1, 1),
    (1, 96, 4, 128, 96, 3, 1, 1), (1, 288, 3, 128, 96, 3, 1, 1),
]


CONV_CASES_1X1 = [   # persistent 1x1 GEMM (conv1.cu): res_conv / qkv / proj shapes, ragged pixel counts, several tiles per CTA
    (2, 192, 8, 8, 576, 1, 1, 0), (1, 192, 16, 16, 192, 1, 1, 0), (1, 96, 24, 40, 48, 1, 1, 0), (3, 48, 16, 16, 96, 1, 1, 0),
    (1, 384, 8, 24, 192, 1, 1, 0), (1, 288, 9, 7, 144, 1, 1, 0), (2, 144, 64, 128, 144, 1, 1, 0), (1, 96, 200, 160, 96, 1, 1, 0),
]
CONV_CASES_1X1_POW2 = [   # the NAFBlock 1x1 shapes (HYB:152-169): powers of two, slices of 32..256, up to 1024 outputs / 512 inputs
    (1, 32, 24, 40, 64, 1, 1, 0), (2, 32, 16, 16, 32, 1, 1, 0), (1, 64, 16, 24, 128, 1, 1, 0), (1, 128, 16, 16, 256, 1, 1, 0),
    (1, 256, 8, 24, 512, 1, 1, 0), (1, 512, 8, 8, 1024, 1, 1, 0), (2, 512, 8, 8, 512, 1, 1, 0), (1, 256, 9, 7, 256, 1, 1, 0),
    (1, 128, 16, 16, 64, 1, 1, 0), (3, 64, 70, 30, 64, 1, 1, 0),
]
CONV_CASES_1X1_GATE = [(1, 32, 24, 40, 64, 1, 1, 0), (2, 64, 16, 24, 128, 1, 1, 0), (1, 128, 9, 7, 256, 1, 1, 0), (3, 32, 64, 64, 64, 1, 1, 0)]
CONV_CASES_1X1_SCALE = [(1, 32, 24, 40, 32, 1, 1, 0), (2, 64, 16, 24, 64, 1, 1, 0), (1, 256, 9, 7, 256, 1, 1, 0), (1, 512, 8, 8, 512, 1, 1, 0),
                        (1, 96, 16, 16, 96, 1, 1, 0)]
CONV_CASES_1X1_CAT = [(1, 384, 16, 16, 192, 1, 1, 0), (1, 192, 32, 32, 96, 1, 1, 0), (1, 96, 24, 40, 48, 1, 1, 0), (1, 288, 16, 24, 144, 1, 1, 0),
                      (1, 192, 16, 16, 48, 1, 1, 0)]
# the UNet's first conv (hook 19, first_conv.cu): B = 1, two input planes, 48 outputs; full tiles, ragged tiles (W % 16 != 0,
# H % 8 != 0), one image-wide tile row, the benched 512x512 plane
CONV_CASES_FIRST = [(1, 2, 64, 64, 48, 3, 1, 1), (1, 2, 40, 56, 48, 3, 1, 1), (1, 2, 13, 200, 48, 3, 1, 1), (1, 2, 8, 136, 48, 3, 1, 1),
                    (1, 2, 512, 512, 48, 3, 1, 1)]
CONV_CASES_1X1_STATS = [(2, 192, 16, 16, 192, 1, 1, 0), (1, 96, 32, 32, 48, 1, 1, 0), (3, 144, 16, 24, 144, 1, 1, 0), (2, 48, 64, 64, 96, 1, 1, 0)]


CONV_CASES_RING = [   # row-ring kernel (conv3r.cu): Cin <= 64, Cout in {48, 96}; ragged segment (H % 32 != 0), several items per CTA
    (1, 48, 8, 128, 48, 3, 1, 1), (2, 48, 40, 256, 48, 3, 1, 1), (1, 48, 33, 128, 96, 3, 1, 1), (1, 64, 70, 128, 48, 3, 1, 1),
    (1, 32, 5, 256, 96, 3, 1, 1), (3, 48, 64, 512, 48, 3, 1, 1), (1, 16, 3, 128, 48, 3, 1, 1), (2, 48, 200, 512, 48, 3, 1, 1),
]
CONV_CASES_STACK = [   # N-stacked row-ring kernel (conv3s.cu): one chunk (16..64 channels), 64 + tail, Cout 48 / 96 (two slices);
    # ragged segments (H % 32 != 0, H = 1, 2, 3), several items per CTA, accumulator-ring wrap (> 10 output rows per CTA)
    (1, 48, 8, 128, 48, 3, 1, 1), (2, 48, 40, 256, 48, 3, 1, 1), (1, 48, 33, 128, 96, 3, 1, 1), (1, 64, 70, 128, 48, 3, 1, 1),
    (1, 32, 5, 256, 48, 3, 1, 1), (3, 48, 64, 512, 48, 3, 1, 1), (1, 16, 3, 128, 48, 3, 1, 1), (2, 48, 200, 512, 48, 3, 1, 1),
    (1, 96, 37, 128, 48, 3, 1, 1), (2, 96, 64, 256, 96, 3, 1, 1), (1, 96, 1, 128, 96, 3, 1, 1), (1, 48, 2, 128, 48, 3, 1, 1),
    (1, 96, 130, 256, 48, 3, 1, 1), (16, 48, 64, 128, 48, 3, 1, 1),
]
CONV_CASES_STACK_CAT = [   # virtual concat of two 48-channel tensors (HYB:383), Cout 48 / 96
    (1, 96, 7, 128, 48, 3, 1, 1), (1, 96, 40, 256, 96, 3, 1, 1), (2, 96, 33, 128, 48, 3, 1, 1), (1, 96, 70, 512, 48, 3, 1, 1),
]
CONV_CASES_W64 = [   # 3x3/s1/p1 on 64-wide maps (conv3w.cu)
    (2, 192, 8, 64, 192, 3, 1, 1), (1, 144, 12, 64, 144, 3, 1, 1), (1, 384, 4, 64, 192, 3, 1, 1), (3, 144, 16, 64, 192, 3, 1, 1),
    (1, 288, 8, 64, 144, 3, 1, 1), (1, 192, 64, 64, 192, 3, 1, 1),
]
CONV_CASES_W64_CAT = [(1, 384, 8, 64, 192, 3, 1, 1), (1, 288, 12, 64, 144, 3, 1, 1), (1, 384, 4, 64, 144, 3, 1, 1)]


def check_conv_fused_gn(mode, impl, cases, seed=9):
    """conv3 with GroupNorm(8)+SiLU of its input applied inside the kernel (hook impl 9 / 10, fixed affine parameters)
    vs F.conv2d(F.silu(F.group_norm(x))) in fp64 on the rounded operands.  Error relative to max|ref|."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    oh = OpHandle(mode)
    out = {}
    dt = torch.bfloat16 if mode == "bf16" else torch.float16
    try:
        for (B, Cin, H, W, Cout, k, s, p) in cases:
            x = (torch.randn(B, Cin, H, W, generator=g) * 1.5 + 0.3).to(DEV)
            w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(DEV)
            b = torch.randn(Cout, generator=g).to(DEV)
            ch = torch.arange(Cin, device=DEV)
            gamma = (1.0 + 0.01 * (ch % 7)).double()
            beta = (0.02 * (ch % 5) - 0.03).double()
            a = F.silu(F.group_norm(x.to(dt).double(), 8, gamma, beta, 1e-5))
            ref = F.conv2d(a.to(dt).double(), w.to(dt).double(), b.double(), stride=s, padding=p).float()
            y = oh.conv2d(x, w, b, k, s, p, impl)
            out[f"{B}x{Cin}x{H}x{W}->{Cout}"] = _rel(y, ref)
    finally:
        oh.close()
    return out


def check_conv_stats(mode, impl, cases, seed=5):
    """GroupNorm partial sums from the conv epilogue vs the sums of the (fp64) reference output."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    oh = OpHandle(mode)
    out = {}
    try:
        for (B, Cin, H, W, Cout, k, s, p) in cases:
            x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
            w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(DEV)
            b = torch.randn(Cout, generator=g).to(DEV)
            dt = torch.bfloat16 if mode == "bf16" else torch.float16
            ref = F.conv2d(x.to(dt).double(), w.to(dt).double(), b.double(), stride=s, padding=p)
            y, st = oh.conv2d_stats(x, w, b, k, s, p, impl)
            r = ref.reshape(B, 8, -1)
            rs = torch.stack([r.sum(-1), (r * r).sum(-1)], dim=-1)
            n = r.shape[-1]
            out[f"{B}x{Cin}x{H}x{W}->{Cout}"] = float(((st - rs).abs() / n).max())    # error of the mean / mean square
    finally:
        oh.close()
    return out


def check_conv(mode, impl, cases, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    oh = OpHandle(mode)
    out = {}
    try:
        for (B, Cin, H, W, Cout, k, s, p) in cases:
            x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
            w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(DEV)
            b = torch.randn(Cout, generator=g).to(DEV)
            if mode != "fp32":   # operands rounded the way the kernel stores them: isolates kernel bugs from rounding
                dt = torch.bfloat16 if mode == "bf16" else torch.float16
                xr, wr = x.to(dt).float(), w.to(dt).float()
            else:
                xr, wr = x, w
            ref = F.conv2d(xr.double(), wr.double(), b.double(), stride=s, padding=p).float()
            y = oh.conv2d(x, w, b, k, s, p, impl)
            out[f"{B}x{Cin}x{H}x{W}->{Cout} k{k}s{s}p{p}"] = _rel(y, ref)
    finally:
        oh.close()
    return out


def check_conv1_naf_epilogues(mode, seed=3):
    """conv1's NAFBlock epilogues: hook 13 = SimpleGate (first half x second half, HYB:119-121) * gamma + input (HYB:165-169),
    hook 14 = (conv + bias) * beta + input (HYB:161); the bias vector doubles as the per-channel scale."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    oh = OpHandle(mode)
    dt = torch.bfloat16 if mode == "bf16" else torch.float16
    out = {}
    try:
        for impl, cases in ((13, CONV_CASES_1X1_GATE), (14, CONV_CASES_1X1_SCALE)):
            for (B, Cin, H, W, Cout, k, s, p) in cases:
                x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
                w = (torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5).to(DEV)
                b = torch.randn(Cout, generator=g).to(DEV)
                xr = x.to(dt).double()
                t = F.conv2d(xr, w.to(dt).double(), b.double())
                if impl == 13:
                    t = t[:, :Cout // 2] * t[:, Cout // 2:]
                ref = (t * b[:t.shape[1]].double().view(1, -1, 1, 1) + xr).float()
                y = torch.empty_like(ref)
                _lib.check(oh.lib.xrd_op_conv2d(oh.h, impl, _p(x), _p(w), _p(b), _p(y), B, Cin, H, W, Cout, 1, 1, 0, None))
                torch.cuda.synchronize()
                out[f"hook{impl} {B}x{Cin}x{H}x{W}->{Cout}"] = _rel(y, ref)
    finally:
        oh.close()
    return out


def check_groupnorm(mode):
    g = torch.Generator().manual_seed(1)
    oh = OpHandle(mode)
    out = {}
    try:
        for (B, Cc, H, W, G, act) in [(2, 48, 32, 32, 8, 1), (1, 384, 8, 8, 8, 1), (2, 288, 16, 16, 8, 0), (1, 24, 20, 12, 4, 2),
                                      (1, 32, 16, 16, 8, 2)]:
            x = (torch.randn(B, Cc, H, W, generator=g) * 2 + 0.5).to(DEV)
            ga, be = torch.randn(Cc, generator=g).to(DEV), torch.randn(Cc, generator=g).to(DEV)
            xr = x if mode == "fp32" else x.to(torch.bfloat16 if mode == "bf16" else torch.float16).float()
            ref = F.group_norm(xr.double(), G, ga.double(), be.double(), 1e-5)
            ref = {0: ref, 1: F.silu(ref), 2: F.gelu(ref)}[act].float()
            y = oh.groupnorm_act(x, ga, be, G, act)
            out[f"{B}x{Cc}x{H}x{W} g{G} act{act}"] = float((y - ref).abs().max())
    finally:
        oh.close()
    return out


def check_attention(mode, impl):
    g = torch.Generator().manual_seed(2)
    oh = OpHandle(mode)
    out = {}
    try:
        # last two cases: peaked softmax whose row maximum keeps growing along the key axis -- exercises the in-TMEM
        # rescale of the running output (the reference maximum is only raised when a tile exceeds it by 2^8)
        for (B, heads, d, H, W, peaked) in [(2, 2, 96, 8, 8, 0), (1, 2, 96, 16, 16, 0), (1, 2, 96, 4, 4, 0), (1, 2, 96, 32, 32, 0),
                                            (1, 1, 64, 12, 10, 0), (1, 2, 96, 32, 32, 1), (2, 2, 96, 24, 24, 1),
                                            # production size of the 512x512 configurations: 64x64 = 4096 tokens (64 key tiles)
                                            (1, 2, 96, 64, 64, 0), (2, 2, 96, 64, 64, 1),
                                            # a single image splits every Q tile's key range over two CTAs (attn_tc.cu: nsplit;
                                            # also the 32x32 cases above): peaked, so the two ranges' maxima differ by far
                                            (1, 2, 96, 64, 64, 1),
                                            # launches of more than 74 (image, head, query-pair) items keep two Q tiles per CTA, smaller
                                            # ones run one (attn_tc.cu: ntq): 10 x 2 x 4 = 80 items and 20 x 2 x 3 ragged = 120
                                            (10, 2, 96, 32, 32, 1), (20, 2, 96, 24, 24, 0)]:
            qkv = torch.randn(B, 3 * heads * d, H, W, generator=g)
            if peaked:
                ramp = torch.linspace(0.2, 4.0, H * W).reshape(1, 1, H, W)
                qkv[:, :heads * d] *= 2.0                      # q
                qkv[:, heads * d:2 * heads * d] *= ramp        # k: later keys score (much) higher or lower
            qkv = qkv.to(DEV)
            qr = qkv if mode == "fp32" else qkv.to(torch.bfloat16 if mode == "bf16" else torch.float16).float()
            t = qr.double().reshape(B, 3, heads, d, H * W)
            q, k, v = t[:, 0], t[:, 1], t[:, 2]
            att = torch.softmax(torch.matmul(q.transpose(-2, -1), k) * d ** -0.5, dim=-1)
            ref = torch.matmul(att, v.transpose(-2, -1)).transpose(-2, -1).reshape(B, heads * d, H, W).float()
            y = oh.attention(qkv, heads, d, impl)
            err = float((y - ref).abs().max())
            if peaked:      # near one-hot attention: outputs are single V rows (|v| up to ~4), so compare relative to max|ref|
                err /= float(ref.abs().max())
            out[f"{B}x{heads}x{d}x{H}x{W}" + ("p" if peaked else "")] = err
    finally:
        oh.close()
    return out


# ------------------------------------------------------------------ network-level
def _hybrid(mode):
    m, sd = seeded_state_dict("hybrid")
    m = m.to(DEV).eval()
    m.set_native_mode(mode)
    for sub in (m.nafnet, m.diffusion_unet, m.router, m.fusion):
        sub.set_native_mode(mode)
    return m, {k: v.detach().cpu() for k, v in m.state_dict().items()}


def check_nafnet(mode):
    m, sd = seeded_state_dict("nafnet")
    m = m.to(DEV).set_native_mode(mode)
    out = {}
    g = load_golden("nafnet_256_b1.npz")
    y = m(g["noisy"].to(DEV)).cpu()
    out["golden256"] = float((y - g["out"]).abs().max())
    g = load_golden("nafnet_40x56_b2.npz")
    y = m(g["noisy"].to(DEV)).cpu()
    out["golden40x56"] = float((y - g["out"]).abs().max())
    _, noisy = O.synthetic_xray(3, 64, 96, seed=11)
    ref = O.nafnet_forward({k: v.cpu() for k, v in m.state_dict().items()}, noisy)
    out["oracle64x96_b3"] = float((m(noisy.to(DEV)).cpu() - ref).abs().max())
    return out


def check_unet_teacher(mode):
    """Per-step eps with the reference's own x fed in (teacher forcing) + the fused sampler update."""
    m, sd = _hybrid(mode)
    g = load_golden("hybrid_64_b2_s50.npz")
    out = {}
    ts = O.ddim_timesteps(50, 50)
    worst = 0.0
    for j, n in enumerate(g["keep"].tolist()):
        t = torch.full((2,), ts[n], dtype=torch.long, device=DEV)
        eps = m.diffusion_unet(g["x_in"][j].to(DEV), g["noisy"].to(DEV), t).cpu()
        e = float((eps - g["eps"][j]).abs().max())
        out[f"eps@eval{n}"] = e
        worst = max(worst, e)
    out["eps_worst"] = worst
    out["eps_ref_absmax"] = float(g["eps"].abs().max())
    return out


def check_ddim_standalone(mode):
    m, sd = seeded_state_dict("unet")
    m = m.to(DEV).set_native_mode(mode)
    g = load_golden("ddim_32_b1_s8.npz")
    w = xrd_b200.DiffusionDenoiser(m, noise_steps=50)
    out = {}
    # teacher forced trace through the sampler entry point
    y, eps_tr, xin_tr = w.denoise(g["noisy"].to(DEV), inference_steps=8, return_trace=True, teacher_x=g["x_in"])
    out["n_evals"] = int(eps_tr.shape[0])
    out["teacher_eps_worst"] = float((eps_tr.cpu()[:, :, None] - g["eps"]).abs().max())
    # free running, eager with trace
    y, eps_tr, xin_tr = w.denoise(g["noisy"].to(DEV), inference_steps=8, return_trace=True)
    out["free_eps_worst"] = float((eps_tr.cpu()[:, :, None] - g["eps"]).abs().max())
    out["free_final"] = float((y.cpu() - g["out"]).abs().max())
    # graph path (default) must equal the eager path bit for bit
    y2 = w.denoise(g["noisy"].to(DEV), inference_steps=8)
    y3 = w.denoise(g["noisy"].to(DEV), inference_steps=8)
    out["graph_vs_eager"] = float((y2 - y).abs().max())
    out["graph_replay_stable"] = float((y3 - y2).abs().max())
    return out


def check_router_fusion(mode):
    m, sd = _hybrid(mode)
    g = load_golden("hybrid_64_b2_s50.npz")
    out = {}
    mask = torch.clamp(m.router(g["noisy"].to(DEV)), 0, 1).cpu()
    out["mask_maxabs"] = float((mask - g["mask"]).abs().max())
    q, qr = (mask * 255).to(torch.uint8), (g["mask"] * 255).to(torch.uint8)       # RUN:145 quantisation
    out["mask_u8_mismatch"] = int((q != qr).sum())
    out["mask_gt05_mismatch"] = int(((mask > 0.5) != (g["mask"] > 0.5)).sum())
    out["mask_pixels"] = int(mask.numel())
    fused = m.fusion(g["naf"].to(DEV), g["diff"].to(DEV), g["mask"].to(DEV)).cpu()
    out["fusion_maxabs"] = float((fused - g["fused"]).abs().max())
    return out


def check_hybrid(mode):
    m, sd = _hybrid(mode)
    m.inference_diffusion_steps = 50
    g = load_golden("hybrid_64_b2_s50.npz")
    fused, parts = m(g["noisy"].to(DEV), return_parts=True)
    out = {k + "_maxabs": float((parts[k].cpu() - g[k]).abs().max()) for k in ("naf", "diff", "mask")}
    out["fused_maxabs"] = float((fused.cpu() - g["fused"]).abs().max())
    out["diff_at_clamp0_frac"] = float((g["diff"] == 0).float().mean())
    return out


def check_hybrid_oracle_128(mode):
    """A size the golden files do not cover: CUDA vs the oracle run on this box's CPU."""
    m, sd = _hybrid(mode)
    m.inference_diffusion_steps = 10
    _, noisy = O.synthetic_xray(2, 128, 128, seed=5)
    parts = {}
    ref = O.hybrid_forward(sd, noisy, 10, 50, parts=parts)
    fused, got = m(noisy.to(DEV), return_parts=True)
    out = {k + "_maxabs": float((got[k].cpu() - parts[k]).abs().max()) for k in ("naf", "diff", "mask")}
    out["fused_maxabs"] = float((fused.cpu() - ref).abs().max())
    return out


# ------------------------------------------------------------------ the benched shapes against the oracle
_ORACLE_CACHE = {}


def _oracle_hybrid_512(steps=2):
    """O.hybrid_forward at 512x512, batch 1, `steps` sampler steps on this box's CPU (a few seconds), with the eps trace."""
    key = ("hyb512", steps)
    if key not in _ORACLE_CACHE:
        _, sd = seeded_state_dict("hybrid")
        sd = {k: v.detach().clone() for k, v in sd.items()}
        _, noisy = O.synthetic_xray(1, 512, 512, seed=31)
        parts, tr = {}, {}
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        with torch.no_grad():
            fused = O.hybrid_forward(sd, noisy, steps, 50, parts=parts)
            O.ddim_denoise(sd, noisy, steps, 50, prefix="diffusion_unet.", trace=tr)
        _ORACLE_CACHE[key] = dict(noisy=noisy, fused=fused, parts=parts, eps=torch.stack(tr["eps"]), x_in=torch.stack(tr["x_in"]))
    return _ORACLE_CACHE[key]


def check_hybrid_512(mode, steps=2):
    """BASELINE configs[2]'s image size through the kernels only that size dispatches to (row-ring conv with the in-place
    GroupNorm+SiLU, 4-row halo tiles, 64-wide conv with 192 outputs, attention over 4096 tokens): hybrid, batch 1,
    inference_steps=2 (t = 25, 0) against the oracle; per-evaluation eps teacher-forced and free-running."""
    ref = _oracle_hybrid_512(steps)
    m, _ = _hybrid(mode)
    m.inference_diffusion_steps = steps
    x = ref["noisy"].to(DEV)
    fused, got = m(x, return_parts=True)
    out = {k + "_maxabs": float((got[k].cpu() - ref["parts"][k]).abs().max()) for k in ("naf", "diff", "mask")}
    out["fused_maxabs"] = float((fused.cpu() - ref["fused"]).abs().max())
    w = m.diffusion_wrapper
    _, eps_t, _ = w.denoise(x, steps, return_trace=True, teacher_x=ref["x_in"])
    out["eps_teacher_worst"] = float((eps_t.cpu()[:, :, None] - ref["eps"]).abs().max())
    y_free, eps_f, _ = w.denoise(x, steps, return_trace=True)
    out["eps_free_worst"] = float((eps_f.cpu()[:, :, None] - ref["eps"]).abs().max())
    out["graph_vs_eager_diff"] = float((got["diff"] - torch.clamp(y_free, 0, 1)).abs().max())
    out["eps_ref_absmax"] = float(ref["eps"].abs().max())
    out["n_evals"] = int(eps_t.shape[0])
    return out


def check_unet_teacher_256(mode, evals=3):
    """configs[1]'s image size: per-evaluation eps at 256x256, batch 1, the first `evals` timesteps of DDIM-50, teacher-forced
    with the oracle's own x (oracle on this box's CPU)."""
    key = ("unet256", evals)
    if key not in _ORACLE_CACHE:
        _, sd = seeded_state_dict("unet")
        sd = {k: v.detach().clone() for k, v in sd.items()}
        _, noisy = O.synthetic_xray(1, 256, 256, seed=33)
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        xs, es = [], []
        x = noisy.clone()
        _, alpha, alpha_hat = O.ddim_tables(50)
        with torch.no_grad():
            for i in O.ddim_timesteps(50, 50)[:evals]:
                e = O.unet_forward(sd, x, noisy, torch.full((1,), i, dtype=torch.long))
                xs.append(x.clone()); es.append(e)
                x = O.ddim_update(x, e, alpha[i], alpha_hat[i])
        _ORACLE_CACHE[key] = dict(noisy=noisy, x_in=xs, eps=es, ts=O.ddim_timesteps(50, 50)[:evals])
    ref = _ORACLE_CACHE[key]
    m, _ = seeded_state_dict("unet")
    m = m.to(DEV).set_native_mode(mode)
    out, worst = {}, 0.0
    for x, e, t in zip(ref["x_in"], ref["eps"], ref["ts"]):
        got = m(x.to(DEV), ref["noisy"].to(DEV), torch.full((1,), t, dtype=torch.long, device=DEV)).cpu()
        err = float((got - e).abs().max())
        out[f"eps@t{t}"] = err
        worst = max(worst, err)
    out["eps_worst"] = worst
    return out


def check_modes_agree_512_b16(steps=50):
    return check_modes_agree_512(16, steps, 21)


def check_modes_agree_512(batch=16, steps=50, seed=21):
    """configs[2] as benched (512x512, batch 16, DDIM-50, the CUDA-graph path with micro-batching): the default f16 mode against
    the fp32 check mode of the same library, which check_hybrid_512 / the goldens pin to the oracle."""
    m, _ = _hybrid("fp16")
    m.inference_diffusion_steps = steps
    _, noisy = O.synthetic_xray(batch, 512, 512, seed=seed)
    x = noisy.to(DEV)
    y16, p16 = m(x, return_parts=True)
    m.set_native_mode("fp32")
    y32, p32 = m(x, return_parts=True)
    out = {k + "_maxabs": float((p16[k] - p32[k]).abs().max()) for k in ("naf", "diff", "mask")}
    out["fused_maxabs"] = float((y16 - y32).abs().max())
    d = (p16["diff"] - p32["diff"]).abs()
    out["diff_rms"] = float(d.pow(2).mean().sqrt())
    out["diff_frac_over_1e-2"] = float((d > 1e-2).float().mean())
    out["diff_per_image_maxabs"] = [round(float(v), 5) for v in d.flatten(1).max(dim=1).values]
    out["pixels"] = int(d.numel())
    out["diff_at_clamp0_frac"] = float((p32["diff"] == 0).float().mean())
    return out


def check_reference_tf32_noise(steps=50, batch=2, size=512):
    """How far the reference AS SHIPPED is from its own fp32 arithmetic on a GPU: it enables TF32 for cuDNN and cuBLAS at import
    (HYB:30-31, DDIM:18-19, NAF:34-35), i.e. contractions with 10-bit mantissas -- the mantissa width of this library's f16
    operands.  The oracle (the same ATen ops) is run on the GPU twice, TF32 allowed / forbidden, free-running DDIM-`steps`; the
    difference is the reference's own numerical noise at this configuration, the yardstick for the 16-bit mode's deviation
    measured by check_modes_agree_512_b16."""
    _, sd = seeded_state_dict("hybrid")
    sd = {k: v.detach().to(DEV) for k, v in sd.items()}
    _, noisy = O.synthetic_xray(batch, size, size, seed=21)
    x = noisy.to(DEV)
    res = {}
    outs = {}
    with torch.no_grad():
        for name, allow in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = allow
            torch.backends.cuda.matmul.allow_tf32 = allow
            tr = {}
            outs[name] = (O.ddim_denoise(sd, x, steps, 50, prefix="diffusion_unet.", trace=tr), torch.stack(tr["eps"]))
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    d = (outs["tf32"][0] - outs["fp32"][0]).abs()
    res["final_maxabs"] = float(d.max())
    res["final_rms"] = float(d.pow(2).mean().sqrt())
    res["final_frac_over_1e-2"] = float((d > 1e-2).float().mean())
    res["eps_first_step_maxabs"] = float((outs["tf32"][1][0] - outs["fp32"][1][0]).abs().max())
    res["eps_free_worst"] = float((outs["tf32"][1] - outs["fp32"][1]).abs().max())
    res["pixels"] = int(d.numel())
    return res


def check_expert(mode):
    """ExpertDenoiser against vectors of the unmodified reference class (tests/golden/make_golden_expert.py) and the oracle."""
    g = load_golden("expert_64_b2.npz")
    m, sd = seeded_state_dict("expert")
    m = m.to(DEV).set_native_mode(mode)
    out = {}
    out["golden64_b2"] = float((m(g["noisy"].to(DEV)).cpu() - g["out"]).abs().max())
    out["golden40x56"] = float((m(g["noisy_r"].to(DEV)).cpu() - g["out_r"]).abs().max())
    m32, _ = seeded_state_dict("expert32")
    m32 = m32.to(DEV).set_native_mode(mode)
    out["golden40x56_base32"] = float((m32(g["noisy_r"].to(DEV)).cpu() - g["out_r_base32"]).abs().max())
    out["ref_absmax"] = float(g["out"].abs().max())
    # the served shape (RUN:197-201: every upload is resized to 512x512), oracle on this box's CPU
    _, noisy = O.synthetic_xray(1, 512, 512, seed=35)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref = O.expert_forward({k: v.detach().cpu() for k, v in m.state_dict().items()}, noisy)
    out["oracle512"] = float((m(noisy.to(DEV)).cpu() - ref).abs().max())
    return out


def check_f16_range(mode="fp16"):
    """The f16 mode's range contract: activations are stored with saturation and the range audit (xrd_set_range_audit) counts
    every clipped or non-finite element.  (1) seeded weights: nothing saturates and the largest magnitude is far from 65504;
    (2) a UNet whose first conv is scaled so that the residual stream of level 0 passes 6e4: the audit must report it and
    check_range() must raise; the fp32 mode of the same weights still matches the oracle."""
    out = {}
    m, _ = _hybrid(mode)
    m.inference_diffusion_steps = 5
    m.set_range_audit(True)
    _, noisy = O.synthetic_xray(1, 128, 128, seed=41)
    m(noisy.to(DEV))
    r = m.range_report()
    out.update({"normal_" + k: v for k, v in r.items()})
    # stressed weights
    u, sd = seeded_state_dict("unet")
    with torch.no_grad():
        u.in_conv.weight.mul_(3.0e5)
        u.in_conv.bias.mul_(3.0e5)
    u = u.to(DEV).set_native_mode(mode)
    u.set_range_audit(True)
    _, noisy = O.synthetic_xray(1, 64, 64, seed=42)
    t = torch.full((1,), 25, dtype=torch.long, device=DEV)
    eps16 = u(noisy.to(DEV), noisy.to(DEV), t).cpu()
    r = u.range_report(reset=False)
    out.update({"stress_" + k: v for k, v in r.items()})
    try:
        u.check_range()
        out["stress_raised"] = False
    except xrd_b200.XrdError:
        out["stress_raised"] = True
    u.set_range_audit(False).set_native_mode("fp32")
    eps32 = u(noisy.to(DEV), noisy.to(DEV), t).cpu()
    with torch.no_grad():
        ref = O.unet_forward({k: v.detach().cpu() for k, v in u.state_dict().items()}, noisy, noisy, t.cpu())
    out["stress_fp32_vs_oracle_rel"] = _rel(eps32, ref)
    out["stress_f16_vs_oracle_rel"] = _rel(eps16, ref)
    out["stress_f16_finite"] = bool(torch.isfinite(eps16).all())
    return out


def check_nan_propagation(mode):
    """A NaN in the input must travel the way it does in the reference: through every network to the final nan_to_num
    (HYB:615-624), not be laundered into a finite number by a saturating store or an fminf/fmaxf clamp."""
    m, sd = _hybrid(mode)
    m.inference_diffusion_steps = 2
    _, noisy = O.synthetic_xray(1, 64, 64, seed=43)
    noisy[0, 0, 10, 20] = float("nan")
    parts = {}
    with torch.no_grad():
        ref = O.hybrid_forward(sd, noisy, 2, 50, parts=parts)
    fused, got = m(noisy.to(DEV), return_parts=True)
    out = {"ref_" + k + "_zero_frac": float((parts[k] == 0).float().mean()) for k in ("naf", "diff", "mask")}
    out.update({k + "_zero_frac": float((got[k] == 0).float().mean()) for k in ("naf", "diff", "mask")})
    out["fused_maxabs"] = float((fused.cpu() - ref).abs().max()) if torch.isfinite(ref).all() else float("nan")
    out["fused_finite"] = bool(torch.isfinite(fused).all())
    w = m.diffusion_wrapper
    _, eps, _ = w.denoise(noisy.to(DEV), 2, return_trace=True)
    out["eps_nan_frac"] = float(torch.isnan(eps).float().mean())
    return out


CHECKS = {
    "conv_simt_fp32": lambda: check_conv("fp32", 0, CONV_CASES_SIMT),
    "conv_simt_bf16": lambda: check_conv("bf16", 0, CONV_CASES_SIMT),
    "groupnorm_fp32": lambda: check_groupnorm("fp32"),
    "groupnorm_bf16": lambda: check_groupnorm("bf16"),
    "attention_simt_fp32": lambda: check_attention("fp32", 0),
    "conv_tc_bf16": lambda: check_conv("bf16", 1, CONV_CASES_TC),
    "conv_tc_fp16": lambda: check_conv("fp16", 1, CONV_CASES_TC),
    "conv_halo_fp16": lambda: check_conv("fp16", 2, CONV_CASES_HALO),
    "conv_halo_bf16": lambda: check_conv("bf16", 2, CONV_CASES_HALO),
    "conv_halo_cat_fp16": lambda: check_conv("fp16", 3, CONV_CASES_HALO_CAT),
    "conv_tc_cat_fp16": lambda: check_conv("fp16", 4, CONV_CASES_HALO_CAT),
    "conv_halo_stats_fp16": lambda: check_conv_stats("fp16", 2, CONV_CASES_HALO),
    "conv_halo_gn_fp16": lambda: check_conv_fused_gn("fp16", 9, CONV_CASES_HALO),
    "conv_halo_gn_cat_fp16": lambda: check_conv_fused_gn("fp16", 10, CONV_CASES_HALO_CAT),
    "conv_halo_gn_bf16": lambda: check_conv_fused_gn("bf16", 9, CONV_CASES_HALO),
    "conv3r_fp16": lambda: check_conv("fp16", 11, CONV_CASES_RING),
    "conv3r_bf16": lambda: check_conv("bf16", 11, CONV_CASES_RING),
    "conv3r_gn_fp16": lambda: check_conv_fused_gn("fp16", 12, CONV_CASES_RING),
    "conv3r_stats_fp16": lambda: check_conv_stats("fp16", 11, CONV_CASES_RING),
    "conv3s_fp16": lambda: check_conv("fp16", 15, CONV_CASES_STACK),
    "conv3s_bf16": lambda: check_conv("bf16", 15, CONV_CASES_STACK),
    "conv3s_cat_fp16": lambda: check_conv("fp16", 17, CONV_CASES_STACK_CAT),
    "conv3s_stats_fp16": lambda: check_conv_stats("fp16", 15, CONV_CASES_STACK),
    "conv3s_cat_stats_fp16": lambda: check_conv_stats("fp16", 17, CONV_CASES_STACK_CAT),
    "conv3s_gn_fp16": lambda: check_conv_fused_gn("fp16", 16, CONV_CASES_STACK),
    "conv3s_gn_cat_fp16": lambda: check_conv_fused_gn("fp16", 18, CONV_CASES_STACK_CAT),
    "conv3w_fp16": lambda: check_conv("fp16", 7, CONV_CASES_W64),
    "conv3w_bf16": lambda: check_conv("bf16", 7, CONV_CASES_W64),
    "conv3w_cat_fp16": lambda: check_conv("fp16", 8, CONV_CASES_W64_CAT),
    "conv3w_stats_fp16": lambda: check_conv_stats("fp16", 7, CONV_CASES_W64),
    "conv1_fp16": lambda: check_conv("fp16", 5, CONV_CASES_1X1),
    "conv1_bf16": lambda: check_conv("bf16", 5, CONV_CASES_1X1),
    "conv1_pow2_fp16": lambda: check_conv("fp16", 5, CONV_CASES_1X1_POW2),
    "conv1_naf_epilogues_fp16": lambda: check_conv1_naf_epilogues("fp16"),
    "conv1_cat_fp16": lambda: check_conv("fp16", 6, CONV_CASES_1X1_CAT),
    "conv1_stats_fp16": lambda: check_conv_stats("fp16", 5, CONV_CASES_1X1_STATS),
    "first_conv_fp16": lambda: check_conv("fp16", 19, CONV_CASES_FIRST),
    "first_conv_bf16": lambda: check_conv("bf16", 19, CONV_CASES_FIRST),
    "first_conv_stats_fp16": lambda: check_conv_stats("fp16", 19, CONV_CASES_FIRST),
    "attention_tc_bf16": lambda: check_attention("bf16", 1),
    "attention_tc_fp16": lambda: check_attention("fp16", 1),
    "nafnet_fp32": lambda: check_nafnet("fp32"),
    "unet_teacher_fp32": lambda: check_unet_teacher("fp32"),
    "ddim_fp32": lambda: check_ddim_standalone("fp32"),
    "router_fusion_fp32": lambda: check_router_fusion("fp32"),
    "hybrid_fp32": lambda: check_hybrid("fp32"),
    "nafnet_bf16": lambda: check_nafnet("bf16"),
    "unet_teacher_bf16": lambda: check_unet_teacher("bf16"),
    "ddim_bf16": lambda: check_ddim_standalone("bf16"),
    "hybrid_bf16": lambda: check_hybrid("bf16"),
    "unet_teacher_fp16": lambda: check_unet_teacher("fp16"),
    "hybrid_fp16": lambda: check_hybrid("fp16"),
    "hybrid128_fp32": lambda: check_hybrid_oracle_128("fp32"),
    "hybrid128_bf16": lambda: check_hybrid_oracle_128("bf16"),
    "hybrid512_fp32": lambda: check_hybrid_512("fp32"),
    "hybrid512_fp16": lambda: check_hybrid_512("fp16"),
    "hybrid512_bf16": lambda: check_hybrid_512("bf16"),
    "unet256_fp32": lambda: check_unet_teacher_256("fp32"),
    "unet256_fp16": lambda: check_unet_teacher_256("fp16"),
    "modes_512_b16": lambda: check_modes_agree_512_b16(),
    "ref_tf32_noise": lambda: check_reference_tf32_noise(),
    "expert_fp32": lambda: check_expert("fp32"),
    "expert_fp16": lambda: check_expert("fp16"),
    "expert_bf16": lambda: check_expert("bf16"),
    "f16_range": lambda: check_f16_range(),
    "nan_fp16": lambda: check_nan_propagation("fp16"),
    "nan_fp32": lambda: check_nan_propagation("fp32"),
}

if __name__ == "__main__":
    import json
    name = sys.argv[1]
    res = CHECKS[name]()
    print("RESULT " + json.dumps({name: res}))
