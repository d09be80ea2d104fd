"""In-situ launch profile of the sampler loop (eager launches, two CUDA events around every launch: xrd_profile_begin/_end) at a
given batch: where one evaluation goes at the clocks of the real kernel mix.  usage: insitu_profile.py B SIZE STEPS"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xrd_b200  # noqa: E402
from xrd_b200 import _lib  # noqa: E402
from synthetic_data import synthetic_xray, randomize_identity_params  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
torch.manual_seed(1234)
m = xrd_b200.HybridDenoisingRouter({}, {}, inference_diffusion_steps=steps).eval()
randomize_identity_params(m.state_dict(), 99)
m = m.cuda()
w = m.diffusion_wrapper
_, noisy = synthetic_xray(B, S, S, seed=7)
x = noisy.cuda()
for _ in range(2):
    w.denoise(x, steps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); w.denoise(x, steps); e1.record(); torch.cuda.synchronize()
graph_ms = e0.elapsed_time(e1)
m.diffusion_unet.use_cuda_graph = False
m.use_cuda_graph = False
w.denoise(x, steps)
torch.cuda.synchronize()
_lib.profile_begin()
w.denoise(x, steps)
prof = _lib.profile_end()
tot = sum(v[1] for v in prof.values())
nl = sum(v[0] for v in prof.values())
print(f"sampler B={B} {S}x{S} steps={steps}: graph replay {graph_ms:.2f} ms; eager in-situ kernel time {tot:.2f} ms over {nl} launches")
for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms  {n:5d} x {1e3 * ms / n:8.1f} us  {100 * ms / tot:5.1f} %  {k[:110]}")
