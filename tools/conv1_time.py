"""Time the persistent 1x1 GEMM (impl 5) next to the per-tap kernel (impl 1) at the UNet's 1x1 shapes (batch 16)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
B = 16
for (cin, hw, cout) in [(192, 64, 576), (192, 64, 192), (384, 64, 192), (96, 512, 48), (192, 256, 96), (288, 128, 144), (48, 256, 96)]:
    x = torch.randn(B, cin, hw, hw, device=DEV)
    w = torch.randn(cout, cin, 1, 1, device=DEV) * 0.05
    b = torch.randn(cout, device=DEV)
    gb = B * hw * hw * (cin + cout) * 2 / 1e9
    for impl in (5, 1):
        oh = OpHandle("fp16")
        oh.conv2d(x, w, b, 1, 1, 0, impl)
        ms = oh.time_last(20)
        print(f"impl={impl} 1x1 {cin}->{cout} @{hw}^2 x{B}: {ms*1e3:.1f} us  {gb/ms:.0f} GB/s algorithmic", flush=True)
        oh.close()
    del x, w
