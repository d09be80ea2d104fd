"""Time the tcgen05 implicit-GEMM conv (and the CUDA-core kernel beside it) at the UNet's real
layer shapes through the C ABI op hook; prints TFLOP/s and writes gpurun_out/conv_sweep.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402

SHAPES = [  # (B, Cin, H, W, Cout, k, stride, pad) -- BASELINE config 3 layer shapes at batch 16
    (16, 48, 512, 512, 48, 3, 1, 1), (16, 96, 256, 256, 96, 3, 1, 1), (16, 144, 128, 128, 144, 3, 1, 1),
    (16, 192, 64, 64, 192, 3, 1, 1), (16, 384, 64, 64, 192, 3, 1, 1), (16, 192, 256, 256, 96, 3, 1, 1),
    (16, 96, 512, 512, 48, 3, 1, 1), (16, 192, 64, 64, 576, 1, 1, 0), (16, 96, 256, 256, 96, 3, 2, 1),
]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    impls = [1, 0] if "--simt" in sys.argv else ([2] if "--halo" in sys.argv else [1])
    res = []
    for (B, Cin, H, W, Cout, k, s, p) in SHAPES:
        x = torch.randn(B, Cin, H, W, device=DEV)
        w = torch.randn(Cout, Cin, k, k, device=DEV) * 0.05
        b = torch.randn(Cout, device=DEV)
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        flops = 2.0 * B * Ho * Wo * Cout * Cin * k * k
        for impl in impls:
            if impl == 2 and (k != 3 or s != 1 or W % 128 != 0):
                continue
            oh = OpHandle(mode)
            try:
                oh.conv2d(x, w, b, k, s, p, impl)
                ms = oh.time_last(10 if impl else 3)
                tf = flops / ms / 1e9
                res.append(dict(shape=[B, Cin, H, W, Cout, k, s, p], impl=impl, ms=ms, tflops=tf))
                print(f"impl={impl} {B}x{Cin}x{H}x{W}->{Cout} k{k}s{s}: {ms:.3f} ms  {tf:.1f} TFLOP/s", flush=True)
            finally:
                oh.close()
        del x, w
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"conv_sweep_{mode}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
