"""Times the N-stacked row-ring kernel (conv3s.cu) against the kernels it replaces, batch 16, through the C-ABI op hook
(CUDA events, warm clocks).   python tools/conv3s_time.py  -> gpurun_out/conv3s_time.json
impl: 2 conv3 (halo), 11 conv3r, 12 conv3r + fused GN, 15 conv3s, 16 conv3s + fused GN, 17 / 18 the same over a virtual concat."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402

CASES = [  # (Cin, Cout, H, impls)
    (48, 48, 512, (11, 12, 15, 16)), (48, 48, 256, (11, 15, 16)), (48, 96, 256, (11, 15, 16)),
    (96, 48, 512, (2, 15, 16, 17, 18)), (96, 48, 256, (2, 15, 16, 17, 18)),
    (96, 96, 256, (2, 15, 16)), (96, 96, 128, (2, 15, 16)), (48, 48, 128, (2, 15, 16)),
]


def main():
    B = int(os.environ.get("B", "16"))
    oh = G.OpHandle("fp16")
    out = []
    for ci, co, hh, impls in CASES:
        x = torch.randn(B, ci, hh, hh, device=G.DEV)
        w = torch.randn(co, ci, 3, 3, device=G.DEV) * 0.05
        b = torch.zeros(co, device=G.DEV)
        fl = 2.0 * B * hh * hh * co * ci * 9
        row = {"cin": ci, "cout": co, "hw": hh}
        for impl in impls:
            try:
                oh.conv2d(x, w, b, 3, 1, 1, impl)
                ms = oh.time_last(10)
                row[f"impl{impl}_us"] = round(ms * 1e3, 1)
                row[f"impl{impl}_tflops"] = round(fl / ms / 1e9, 1)
            except Exception as e:  # noqa: BLE001
                row[f"impl{impl}_error"] = str(e)[:200]
        print(row, flush=True)
        out.append(row)
        del x, w
    oh.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "conv3s_time.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
