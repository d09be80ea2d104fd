"""NAFNet + router once at batch 16, 512x512 (BASELINE configs[2]) for an `ncu --set full` capture of the HBM-bound kernels:
   ncu --set full --clock-control none --import-source on -k regex:"k_dwconv|k_layernorm|k_conv_simt|k_simple_gate|k_scale_nc|k_conv_tc" \
       --launch-count 40 python tools/ncu_targets_naf.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402
from oracle import xrd_oracle as O  # noqa: E402
m, _ = G._hybrid("fp16")
_, noisy = O.synthetic_xray(16, 512, 512, seed=3)
x = noisy.to(G.DEV)
y = m.nafnet(x)
z = m.router(x)
torch.cuda.synchronize()
print("done", float(y.mean()), float(z.mean()))
