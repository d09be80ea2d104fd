"""Time the 64-wide 3x3 kernel (impl 7) next to the per-tap kernel (impl 1) at the UNet's level-3 shapes (batch 16)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
B = 16
for (cin, cout) in [(192, 192), (384, 192), (144, 144), (288, 144), (144, 192)]:
    x = torch.randn(B, cin, 64, 64, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
    b = torch.randn(cout, device=DEV)
    fl = 2.0 * B * 4096 * cin * cout * 9
    for impl in (7, 1):
        oh = OpHandle("fp16")
        oh.conv2d(x, w, b, 3, 1, 1, impl)
        ms = oh.time_last(20)
        print(f"impl={impl} 3x3 {cin}->{cout} @64^2 x{B}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s", flush=True)
        oh.close()
    del x, w
