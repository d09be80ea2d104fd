"""conv3 with (impl 9) and without (impl 2) GroupNorm+SiLU applied in shared memory, next to the stand-alone
GroupNorm+SiLU pass it would replace, at the UNet's layer shapes (batch 16)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
B = 16
for (cin, hw, cout) in [(48, 512, 48), (48, 256, 48), (48, 256, 96), (48, 128, 48)]:
    x = torch.randn(B, cin, hw, hw, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
    b = torch.randn(cout, device=DEV)
    t = {}
    for impl in (2, 11, 12):
        oh = OpHandle("fp16")
        oh.conv2d(x, w, b, 3, 1, 1, impl)
        t[impl] = oh.time_last(10) * 1e3
        oh.close()
    oh = OpHandle("fp16")
    oh.groupnorm_act(x, torch.ones(cin, device=DEV), torch.zeros(cin, device=DEV), 8, 1)
    tg = oh.time_last(10) * 1e3          # statistics + apply; the apply alone is ~65 % of it
    oh.close()
    print(f"{cin:3d}->{cout:3d} @{hw}: conv3 {t[2]:6.1f} us  conv3r {t[11]:6.1f} us  conv3r+GN {t[12]:6.1f} us  stats+apply pass {tg:6.1f} us", flush=True)
    del x, w
