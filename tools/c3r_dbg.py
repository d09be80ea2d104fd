# NOTE: the in-kernel clock64 traces and the XRD_C3_DBG / XRD_C3R_DBG experiment switches exist only in trace builds:
#   XRD_TRACE=1 XRD_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()"
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
x = torch.randn(16, 48, 512, 512, device=DEV); w = torch.randn(48, 48, 3, 3, device=DEV) * 0.05; b = torch.randn(48, device=DEV)
for dbg in [int(a) for a in sys.argv[1:]] or [0]:
    os.environ["XRD_C3R_DBG"] = str(dbg)
    os.environ["XRD_C3R_PROF"] = "2"
    oh = OpHandle("fp16"); oh.conv2d(x, w, b, 3, 1, 1, 11)
    os.environ["XRD_C3R_PROF"] = "0"
    ms = oh.time_last(10); oh.close()
    print(f"dbg={dbg}: {ms*1e3:.1f} us", flush=True)
