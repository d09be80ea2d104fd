import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
for (c, hw) in [(48, 512), (96, 256), (144, 128), (192, 64), (96, 512)]:
    x = torch.randn(16, c, hw, hw, device=DEV)
    oh = OpHandle("fp16"); oh.groupnorm_act(x, torch.ones(c, device=DEV), torch.zeros(c, device=DEV), 8, 1); ms = oh.time_last(20); oh.close()
    gb = 16 * c * hw * hw * 2 * 3 / 1e9
    print(f"gn stats+apply {c}ch @{hw}: {ms*1e3:.1f} us  ({gb/ms:.2f} TB/s over 2 reads + 1 write)", flush=True)
