"""Soak test of the warp-specialised tcgen05 kernels: hundreds of thousands of back-to-back launches of each variant through the op
hooks (a protocol race shows up as a trapped launch with a tagged 'mbarrier wait timed out' message).   python tools/stress_tc.py [rounds]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(os.environ.get("ITERS", "15000"))
B = 16
cases = [(48, 512, 48, 16), (48, 512, 48, 15), (96, 256, 96, 16), (96, 512, 48, 18), (48, 256, 96, 16), (96, 128, 96, 16), (96, 256, 96, 15)]
total = 0
t0 = time.time()
for r in range(rounds):
    for (cin, hw, cout, impl) in cases:
        oh = OpHandle("fp16")
        x = torch.randn(B, cin, hw, hw, device=DEV)
        w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
        b = torch.randn(cout, device=DEV)
        oh.conv2d(x, w, b, 3, 1, 1, impl)
        ms = oh.time_last(iters)
        total += iters
        print(f"round {r} conv3s impl {impl} {cin}->{cout} @{hw}: {iters} launches, {ms * 1e3:.1f} us each, {total} total, {time.time() - t0:.0f} s", flush=True)
        oh.close(); del x, w
    oh = OpHandle("fp16")
    qkv = torch.randn(B, 3 * 2 * 96, 64, 64, device=DEV)
    oh.attention(qkv, 2, 96, 1)
    ms = oh.time_last(iters // 2)
    print(f"round {r} attention: {iters // 2} launches, {ms * 1e3:.1f} us each", flush=True)
    oh.close()
print("soak ok", total)
