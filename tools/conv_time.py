"""Live timings (CUDA events through the C-ABI op hook) of the UNet's main conv shapes at batch 16:
   python tools/conv_time.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
B = 16
CASES = [  # (impl, cin, hw, cout, name)
    (11, 48, 512, 48, "conv3r 48->48 @512"), (2, 96, 512, 48, "conv3 96->48 @512 (streamed, 4-row)"), (2, 96, 256, 96, "conv3 96->96 @256 (streamed)"),
    (2, 192, 256, 96, "conv3 192->96 @256"), (2, 144, 128, 144, "conv3 144->144 @128"), (2, 288, 128, 144, "conv3 288->144 @128"),
    (7, 192, 64, 192, "conv3w 192->192 @64"), (7, 384, 64, 192, "conv3w 384->192 @64"), (11, 48, 256, 96, "conv3r 48->96 @256"),
]
for (impl, cin, hw, cout, name) in CASES:
    x = torch.randn(B, cin, hw, hw, device=DEV); w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05; b = torch.randn(cout, device=DEV)
    oh = OpHandle("fp16")
    oh.conv2d(x, w, b, 3, 1, 1, impl)
    ms = oh.time_last(20)
    oh.close()
    fl = 2.0 * B * hw * hw * cout * cin * 9
    print(f"{name:40s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
    del x, w
