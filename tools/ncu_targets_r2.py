"""One launch each of the round-2 hot kernels at their batch-16 shapes, for `ncu --set full` captures:
conv3s 48->48 @512 plain and with fused GroupNorm+SiLU, 96->96 @256 with GN (two chunks, two slices), cat 48+48->48 @512 with GN,
attention at 4096 tokens."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
B = 16
oh = OpHandle("fp16")
for (cin, hw, cout, impl) in [(48, 512, 48, 15), (48, 512, 48, 16), (96, 256, 96, 16), (96, 512, 48, 18)]:
    x = torch.randn(B, cin, hw, hw, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
    b = torch.randn(cout, device=DEV)
    oh.conv2d(x, w, b, 3, 1, 1, impl)
    del x, w
qkv = torch.randn(B, 3 * 2 * 96, 64, 64, device=DEV)
oh.attention(qkv, 2, 96, 1)
torch.cuda.synchronize()
print("done")
