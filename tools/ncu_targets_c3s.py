"""One launch each of the N-stacked row-ring kernel at its main shapes (batch 16) for `ncu --set full` captures."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
B = 16
oh = OpHandle("fp16")
for (cin, hw, cout, impl) in [(48, 512, 48, 15), (48, 512, 48, 16), (96, 512, 48, 17), (96, 256, 96, 15)]:
    x = torch.randn(B, cin, hw, hw, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
    b = torch.randn(cout, device=DEV)
    oh.conv2d(x, w, b, 3, 1, 1, impl)
    del x, w
torch.cuda.synchronize()
print("done")
