import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
B = 16
for (cin, hw, cout, impl) in [(96, 256, 96, 2), (96, 512, 48, 2), (192, 256, 96, 2), (144, 128, 144, 2), (288, 128, 144, 2), (192, 64, 192, 7), (384, 64, 192, 7)]:
    x = torch.randn(B, cin, hw, hw, device=DEV); w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05; b = torch.randn(cout, device=DEV)
    fl = 2.0 * B * hw * hw * cin * cout * 9
    r = []
    for wb in ("0", "1"):
        os.environ["XRD_C3_TALL"] = wb
        oh = OpHandle("fp16"); oh.conv2d(x, w, b, 3, 1, 1, impl); ms = oh.time_last(10); oh.close()
        r.append(ms)
    print(f"{cin}->{cout} @{hw} impl={impl}: 2-row tiles {r[0]*1e3:.1f} us, tall tiles {r[1]*1e3:.1f} us ({fl/r[1]/1e9:.0f} TF/s)", flush=True)
    del x, w
