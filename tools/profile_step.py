"""Small fixed workload for ncu: the hybrid at 512x512, batch B, with 2 UNet evaluations
(inference_diffusion_steps=2 -> timesteps 25, 0) so that one call shows every kernel of the path
in its real shape.  Two warm-up calls, then one measured call bracketed by CUDA events."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xrd_b200  # noqa: E402
from oracle import xrd_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
what = sys.argv[4] if len(sys.argv) > 4 else "hybrid"      # "hybrid" | "ddim" (sampler loop only) | "naf" | "router"
torch.manual_seed(1234)
m = xrd_b200.HybridDenoisingRouter({}, {}, inference_diffusion_steps=steps).eval()
O.randomize_identity_params(m.state_dict(), 99)
m = m.cuda()
m.use_cuda_graph = os.environ.get("XRD_GRAPH", "1") == "1"
m.diffusion_unet.use_cuda_graph = m.use_cuda_graph
if what == "ddim":
    _w, _steps = m.diffusion_wrapper, steps
    m = lambda x: _w.denoise(x, _steps)   # noqa: E731
elif what == "naf":
    m = m.nafnet
elif what == "router":
    m = m.router
_, noisy = O.synthetic_xray(B, S, S, seed=7)
x = noisy.cuda()
for _ in range(2):
    m(x)
torch.cuda.synchronize()
n0 = xrd_b200.native_kernel_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y = m(x)
e1.record()
torch.cuda.synchronize()
print(f"{what} B={B} {S}x{S} steps={steps}: {e0.elapsed_time(e1):.2f} ms, {xrd_b200.native_kernel_launches() - n0} launches, "
      f"finite={bool(torch.isfinite(y).all())}")
