"""Time the tcgen05 attention kernel at the UNet's shape (B=16, 2 heads x d=96, 64x64 tokens)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
DBG = [int(a) for a in sys.argv[1:]] or [0]
for (B, heads, d, H, W) in [(16, 2, 96, 64, 64), (16, 2, 96, 32, 32)]:
    qkv = torch.randn(B, 3 * heads * d, H, W, device=DEV)
    for mode, dbg in [("fp16", g) for g in DBG] + [("bf16", 0)]:
        oh = OpHandle(mode)
        oh.attention(qkv, heads, d, 1)
        ms = oh.time_last(20)
        fl = 4.0 * B * heads * (H * W) ** 2 * d
        print(f"attention {mode} dbg={dbg} B={B} heads={heads} d={d} N={H*W}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        oh.close()
