"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference; the GPU box has no copy).
The reference's three model files import skimage/matplotlib at module top for their
training code; neither is installed, so two empty stand-in modules are injected
(SURVEY Appendix A).  Nothing of the reference is copied: it is imported, executed
and its outputs stored.

    python tests/golden/make_golden.py

Seeds: weights 1234 (set AFTER import -- importing re-seeds to 42,
HYB:21), NAFBlock beta/gamma + norm affine override 99, inputs 7.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import xrd_oracle as O  # noqa: E402  (seed helpers + synthetic inputs only)


def import_reference():
    sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.metrics")
    skm.peak_signal_noise_ratio = skm.structural_similarity = lambda *a, **k: 0.0
    sk.metrics = skm
    sys.modules["skimage"], sys.modules["skimage.metrics"] = sk, skm
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    sys.path.insert(0, "/root/reference/Backend")
    import hybrid.hybrid3diffusionspeed as HYB
    import DDIM.DDIMModel as DDIM
    import NafNet.NafnetModel as NAF
    return HYB, DDIM, NAF


def digest(sd) -> dict:
    """Per-tensor fingerprint so the GPU-side tests can prove their seeded weights are the
    reference's without shipping 137 MB: (shape, float64 sum, float64 sum of squares)."""
    out = {}
    for k, v in sd.items():
        d = v.detach().double()
        out[k] = [list(v.shape), float(d.sum()), float((d * d).sum())]
    return out


def trace_denoise(wrapper, noisy, steps, dev):
    """The reference loop (HYB:403-418) re-stated with taps; asserted bit-identical to
    wrapper.denoise below, so the taps are the reference's own intermediates."""
    x = noisy.clone()
    xs, es = [], []
    step = max(1, wrapper.noise_steps // steps)
    for i in reversed(range(0, wrapper.noise_steps, step)):
        t = torch.full((x.shape[0],), i, dtype=torch.long, device=dev)
        e = wrapper.model(x, noisy, t)
        xs.append(x.clone()); es.append(e.clone())
        e = torch.clamp(e, -5, 5)
        a = wrapper.alpha[t][:, None, None, None]
        ah = wrapper.alpha_hat[t][:, None, None, None]
        x = (1 / torch.sqrt(a)) * (x - ((1 - a) / torch.sqrt(1 - ah)) * e)
        x = torch.clamp(x, 0, 1)
    return x, torch.stack(xs), torch.stack(es)


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    HYB, DDIM, NAF = import_reference()
    dev = HYB.device
    meta = {"torch": torch.__version__, "seeds": {"weights": 1234, "override": 99, "inputs": 7}}

    # ---- G1: hybrid, 64x64, batch 2, DDIM-50 ------------------------------------------
    torch.manual_seed(1234)
    hyb = HYB.HybridDenoisingRouter(nafnet_params={}, diffusion_params={}, inference_diffusion_steps=50).eval()
    sd = hyb.state_dict()
    meta["hybrid_init_digest"] = digest(sd)              # BEFORE the override: pure seeded init
    O.randomize_identity_params(sd, 99)                  # in place -> the module's parameters
    meta["hybrid_keys"] = {k: list(v.shape) for k, v in sd.items()}
    _, noisy = O.synthetic_xray(2, 64, 64, seed=7)
    fused = hyb(noisy)
    naf = torch.clamp(torch.nan_to_num(hyb.nafnet(noisy), nan=0.0, posinf=1.0, neginf=0.0), 0, 1)
    mask = torch.clamp(torch.nan_to_num(hyb.router(noisy), nan=0.0, posinf=1.0, neginf=0.0), 0, 1)
    diff_ref = hyb.diffusion_wrapper.denoise(noisy, inference_steps=50)
    diff, xs, es = trace_denoise(hyb.diffusion_wrapper, noisy, 50, dev)
    assert torch.equal(diff, diff_ref), "traced loop is not bit-identical to DiffusionDenoiser.denoise"
    assert torch.equal(hyb.fusion(naf, torch.clamp(diff, 0, 1), mask), fused)
    keep = [0, 1, 2, 24, 48, 49]
    np.savez_compressed(os.path.join(HERE, "hybrid_64_b2_s50.npz"),
                        noisy=noisy.numpy(), fused=fused.numpy(), naf=naf.numpy(), mask=mask.numpy(),
                        diff=diff.numpy(), keep=np.array(keep), x_in=xs[keep].numpy(), eps=es[keep].numpy())

    # ---- G2: standalone DDIM classes, 32x32, batch 1, inference_steps=8 (-> 9 evals) -----
    torch.manual_seed(1234)
    unet = DDIM.UNetDiffusion().eval()
    sdu = unet.state_dict()
    meta["unet_init_digest"] = digest(sdu)
    O.randomize_identity_params(sdu, 99)
    wrap = DDIM.DiffusionDenoiser(unet, noise_steps=50)
    _, noisy2 = O.synthetic_xray(1, 32, 32, seed=7)
    out2 = wrap.denoise(noisy2, inference_steps=8)
    out2b, xs2, es2 = trace_denoise(wrap, noisy2, 8, dev)
    assert torch.equal(out2, out2b)
    assert xs2.shape[0] == 9
    # the HYB copy of the UNet must agree with the DDIM copy (chunked vs unchunked attention)
    unet_h = HYB.UNetDiffusion().eval(); unet_h.load_state_dict(sdu)
    t0 = torch.full((1,), 48, dtype=torch.long)
    meta["ddim_vs_hyb_unet_maxabs"] = float((unet_h(xs2[0], noisy2, t0) - es2[0]).abs().max())
    np.savez_compressed(os.path.join(HERE, "ddim_32_b1_s8.npz"), noisy=noisy2.numpy(), out=out2.numpy(),
                        x_in=xs2.numpy(), eps=es2.numpy())

    # ---- G3: BASELINE config 1 -- standalone NAFNet, 256x256, batch 1 ------------------
    torch.manual_seed(1234)
    nafm = NAF.EnhancedNAFNet().eval()
    sdn = nafm.state_dict()
    meta["nafnet_init_digest"] = digest(sdn)
    O.randomize_identity_params(sdn, 99)
    _, noisy3 = O.synthetic_xray(1, 256, 256, seed=7)
    np.savez_compressed(os.path.join(HERE, "nafnet_256_b1.npz"), noisy=noisy3.numpy().astype(np.float32),
                        out=nafm(noisy3).numpy())
    # ---- G4: ragged size (padding path, NAF:304-309), 40x56, batch 2 -------------------
    _, noisy4 = O.synthetic_xray(2, 40, 56, seed=8)
    np.savez_compressed(os.path.join(HERE, "nafnet_40x56_b2.npz"), noisy=noisy4.numpy(), out=nafm(noisy4).numpy())

    # ---- G5: sampler schedule known answers ---------------------------------
    sched = {}
    for ns, st in [(50, 50), (50, 25), (50, 10), (50, 8), (50, 7), (50, 100), (100, 100), (50, 1), (50, 3)]:
        step = max(1, ns // st)
        sched[f"{ns},{st}"] = list(reversed(range(0, ns, step)))
    meta["schedule"] = sched
    w = HYB.DiffusionDenoiser(unet_h, noise_steps=50)
    meta["alpha"] = [float(v) for v in w.alpha]
    meta["alpha_hat"] = [float(v) for v in w.alpha_hat]

    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f)
    for fn in sorted(os.listdir(HERE)):
        p = os.path.join(HERE, fn)
        print(f"{fn:32s} {os.path.getsize(p)/1e6:8.3f} MB  sha1 {hashlib.sha1(open(p,'rb').read()).hexdigest()[:12]}")


if __name__ == "__main__":
    main()
