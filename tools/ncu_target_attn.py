"""One launch of the tcgen05 attention kernel at the UNet's shape (B=16, 2 heads x d=96, 64x64 tokens) for ncu captures."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
oh = OpHandle("fp16")
qkv = torch.randn(16, 3 * 2 * 96, 64, 64, device=DEV)
oh.attention(qkv, 2, 96, 1)
torch.cuda.synchronize()
print("done")
