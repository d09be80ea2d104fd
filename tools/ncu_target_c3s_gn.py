"""One launch of conv3s 48->48 @512x512, batch 16, with the fused GroupNorm+SiLU transform (the dominant kernel of an evaluation),
for an `ncu --set full --import-source on` capture:   bash tools/ncu_one.sh "k_conv3s" tools/ncu_target_c3s_gn.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
oh = OpHandle("fp16")
x = torch.randn(16, 48, 512, 512, device=DEV)
w = torch.randn(48, 48, 3, 3, device=DEV) * 0.05
b = torch.randn(48, device=DEV)
oh.conv2d(x, w, b, 3, 1, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 16)
torch.cuda.synchronize()
print("done")
