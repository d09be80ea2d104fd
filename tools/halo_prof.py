"""Bottleneck analysis of the persistent halo conv: per-tile clock64 traces of block 0 (stderr) plus timings
under the XRD_C3_DBG experiment masks.   python tools/halo_prof.py [fp16|bf16]"""
# NOTE: the in-kernel clock64 traces and the XRD_C3_DBG / XRD_C3R_DBG experiment switches exist only in trace builds:
#   XRD_TRACE=1 XRD_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()"

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402

SHAPES = [(16, 48, 512, 512, 48), (16, 64, 512, 512, 48), (16, 96, 256, 256, 96), (16, 144, 128, 128, 144), (16, 96, 512, 512, 48), (16, 192, 256, 256, 96)]
mode = sys.argv[1] if len(sys.argv) > 1 else "fp16"
variants = [dict(), dict(XRD_C3_DBG="2"), dict(XRD_C3_DBG="1")]
for (B, Cin, H, W, Cout) in SHAPES:
    x = torch.randn(B, Cin, H, W, device=DEV)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV) * 0.05
    b = torch.randn(Cout, device=DEV)
    flops = 2.0 * B * H * W * Cout * Cin * 9
    for v in variants:
        for k in ("XRD_C3_DBG", "XRD_C3_TH"):
            os.environ.pop(k, None)
        os.environ.update(v)
        oh = OpHandle(mode)
        try:
            os.environ["XRD_C3_PROF"] = "2" if v.get("XRD_C3_DBG", "0") in ("0", "15") else "0"
            oh.conv2d(x, w, b, 3, 1, 1, 2)
            os.environ["XRD_C3_PROF"] = "0"
            ms = oh.time_last(10)
            print(f"{v} {B}x{Cin}x{H}x{W}->{Cout}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
            sys.stderr.flush()
        finally:
            oh.close()
