import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
print("start", flush=True)
for (B, H) in [(1, 32), (1, 64), (2, 64), (16, 64)]:
    qkv = torch.randn(B, 576, H, H, device=DEV)
    oh = OpHandle("fp16")
    t0 = time.time()
    y = oh.attention(qkv, 2, 96, 1)
    torch.cuda.synchronize()
    print(f"B={B} H={H} first call ok {time.time()-t0:.2f}s finite={bool(torch.isfinite(y).all())}", flush=True)
    ms = oh.time_last(5)
    print(f"  time_last {ms:.3f} ms", flush=True)
    oh.close()
