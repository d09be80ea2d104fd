"""SM clock and board power while ONE kernel runs back to back (pynvml sampled every 10 ms), next to its time per launch:
tells a kernel that is slow in cycles from one that runs at a power-capped clock.   python tools/clock_probe.py"""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402
import pynvml  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.on = True
        self.rows = []

    def run(self):
        while self.on:
            self.rows.append((pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(self.h) / 1e3))
            time.sleep(0.01)


def probe(name, fn, timer, secs=1.5):
    fn()
    s = Sampler(); s.start()
    t0 = time.time(); n = 0; ms = []
    while time.time() - t0 < secs:
        ms.append(timer(50)); n += 1
    s.on = False; s.join()
    rows = s.rows[len(s.rows) // 3:]
    clk = sorted(r[0] for r in rows)[len(rows) // 2]
    pw = sorted(r[1] for r in rows)[len(rows) // 2]
    print(f"{name:34s} {1e3 * sorted(ms)[len(ms) // 2]:8.1f} us/launch   SM {clk} MHz   {pw:6.0f} W   ({len(rows)} samples)", flush=True)


def main():
    B = 16
    oh = G.OpHandle("fp16")
    for (cin, hw, cout, impl, nm) in [(48, 512, 48, 15, "conv3s 48->48 @512"), (48, 512, 48, 16, "conv3s+GN 48->48 @512"),
                                      (48, 512, 48, 11, "conv3r 48->48 @512"), (96, 256, 96, 15, "conv3s 96->96 @256"),
                                      (96, 256, 96, 2, "conv3 96->96 @256"), (192, 128, 192, 0, "conv(auto) 192->192 @128")]:
        x = torch.randn(B, cin, hw, hw, device=G.DEV)
        w = torch.randn(cout, cin, 3, 3, device=G.DEV) * 0.05
        b = torch.zeros(cout, device=G.DEV)
        try:
            probe(nm, lambda: oh.conv2d(x, w, b, 3, 1, 1, impl if impl else 1), oh.time_last)
        except Exception as e:  # noqa: BLE001
            print(nm, "error", str(e)[:100])
        del x, w
    a = torch.randn(8192, 8192, device=G.DEV, dtype=torch.bfloat16)
    c = torch.randn(8192, 8192, device=G.DEV, dtype=torch.bfloat16)

    def mm_timer(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n // 5):
            torch.matmul(a, c)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n // 5)
    probe("torch.matmul bf16 8192^3 (cuBLAS)", lambda: torch.matmul(a, c), mm_timer)
    big = torch.empty(1 << 30, device=G.DEV, dtype=torch.bfloat16); dst = torch.empty_like(big)

    def cp_timer(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n // 10):
            dst.copy_(big)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n // 10)
    probe("copy 2 GiB (HBM)", lambda: dst.copy_(big), cp_timer)


if __name__ == "__main__":
    main()
