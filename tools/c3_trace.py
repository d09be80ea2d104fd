# NOTE: the in-kernel clock64 traces and the XRD_C3_DBG / XRD_C3R_DBG experiment switches exist only in trace builds:
#   XRD_TRACE=1 XRD_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()"
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
cin, hw, cout = [int(a) for a in sys.argv[1:4]]
x = torch.randn(16, cin, hw, hw, device=DEV); w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05; b = torch.randn(cout, device=DEV)
os.environ["XRD_C3_PROF"] = "2"
oh = OpHandle("fp16"); oh.conv2d(x, w, b, 3, 1, 1, 2)
os.environ["XRD_C3_PROF"] = "0"
print(f"{oh.time_last(10)*1e3:.1f} us")
