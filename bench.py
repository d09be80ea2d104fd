#!/usr/bin/env python
"""bench.py -- denoised images/sec of the full hybrid (DDIM-50 + NAFNet + router + fusion) at 512x512.

    python bench.py --gpus N --steps K --warmup W              (this repo's libxrd.so path)
    python bench.py --impl reference --gpus N --steps K ...    (the reference algorithm on host CPU cores)

Default workload (BASELINE.json configs[2], the configuration the metric is quoted on): HybridDenoisingRouter
with inference_diffusion_steps=50, noise_steps=50, batch 16 per GPU of synthetic 512x512 grayscale
X-ray fields, random-init weights (seed 1234, NAFBlock beta/gamma + norm affine override seed 99).
A "step" is one forward of that batch.  N>1: one process per GPU (torchrun), images sharded by rank,
no collective inside the sampler loop, one NCCL all_gather of the finished (B,1,H,W) outputs per step
(weak scaling: 16 images per GPU).

Other workloads (explicit flags; the default line is unchanged by them):
    --global-batch 256           BASELINE configs[3]: a FIXED batch sharded by image over the ranks (strong scaling)
    --workload tiled1024         BASELINE configs[4]: 1024x1024, wrapper built with noise_steps=100, inference_steps=100,
                                 512 tiles with 64-pixel halos (9 tiles per image), --batch images per GPU (default 8)

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the
same through the public class API with pinned-host inputs and a host read-back inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GF_PER_IMAGE_512 = 21383.0      # algorithmic 2*MAC of the reference op list, hybrid DDIM-50 @512^2
GF_UNET_EVAL_512 = 424.66       # one UNet evaluation @512^2
GF_REST_512 = 112.55 + 31.31 + 6.13   # NAFNet + router + fusion @512^2
UNET_CONV_GF_512 = 347.34       # conv FLOPs of one UNet evaluation @512^2 per image
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")   # written from an `ncu --set full` capture (tools/ncu_traffic.py)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src="fallback")


def top_kernel_traffic(batch, size):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the longest kernel, from the committed ncu capture of THIS
    tree (profiles/top_kernel_traffic.json: bytes, kernel, capture file, batch, size).  null when there is no capture for the
    benched shape -- never a number typed into this file."""
    try:
        d = json.load(open(TRAFFIC_FILE))
        if d.get("batch") == batch and d.get("size") == size:
            return d
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def workload_spec(args, world):
    """Per-workload constants: image size, evaluations per image, tiles, algorithmic GF per image, labels."""
    if args.workload == "tiled1024":
        tiles = 9                                   # 1024 with 512 tiles and 64-pixel halos: origins 0, 384, 512 per axis
        evals = 100
        gf = tiles * (evals * GF_UNET_EVAL_512 + GF_REST_512)
        return dict(size=1024, tile=512, halo=64, tiles=tiles, evals=evals, noise_steps=100, gf_per_image=gf,
                    metric="denoised images/sec at 1024x1024 (hybrid, DDIM-100, 512 tiles + 64 halo)",
                    label=f"hybrid DDIM-100 (noise_steps=100) + NAFNet + router + fusion over 9 overlap tiles of 512x512 per 1024x1024 "
                          f"image (BASELINE configs[4])")
    s = args.size
    return dict(size=s, tile=None, halo=None, tiles=1, evals=50, noise_steps=50, gf_per_image=GF_PER_IMAGE_512 * (s / 512.0) ** 2,
                metric="denoised images/sec at 512x512 (hybrid, DDIM-50)",
                label=f"hybrid DDIM-50 + NAFNet + router + fusion, {s}x{s} grayscale" +
                      (f", global batch {args.global_batch} sharded by image (BASELINE configs[3])" if args.global_batch
                       else f", batch {args.batch} per GPU (BASELINE configs[2])"))


def build_model(dev, mode, spec):
    import xrd_b200
    import synthetic_data as SY                 # seeded weights/inputs only; the oracle is not imported on the GPU arm
    torch.manual_seed(1234)
    m = xrd_b200.HybridDenoisingRouter({}, {"noise_steps": spec["noise_steps"]}, inference_diffusion_steps=spec["evals"]).eval()
    SY.randomize_identity_params(m.state_dict(), 99)
    m = m.to(dev)
    m.set_native_mode(mode)
    if spec["tile"]:
        return m, (lambda x: xrd_b200.denoise_tiled(m, x, spec["tile"], spec["halo"]))
    return m, m


# ------------------------------------------------------------------------------------------------
# the reference algorithm on host cores (the oracle port; the reference is a Python program with nothing to compile)
# ------------------------------------------------------------------------------------------------
CPU_UNET_EVALS = 12     # UNet evaluations per CPU sample: about 10 s of work on the GPU box's host cores


def cpu_threads():
    """torchrun exports OMP_NUM_THREADS=1; BASELINE.md section 3 asks for every host core."""
    n = int(os.environ.get("XRD_CPU_THREADS", "0")) or (os.cpu_count() or 1)
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_state_dict():
    """Seeded weights with the reference's 912 key/shape table (tests/golden/meta.json, written by running the unmodified
    reference).  Built by the oracle: this leg imports nothing of the product package."""
    from oracle import xrd_oracle as O
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "meta.json")))["hybrid_keys"]
    return O.synthetic_state_dict(shapes, 1234)


def cpu_reference_sample(spec, unet_evals, full_image=False):
    """The reference algorithm (oracle port) on host cores, batch 1 at 512x512: `unet_evals` UNet evaluations + NAFNet + router
    + fusion, extrapolated to the workload's evaluations (and tiles) per image.  Returns a dict with the extrapolated
    images/s, the seconds actually spent, and -- full_image -- one whole un-extrapolated DDIM-50 image."""
    from oracle import xrd_oracle as O
    sd = cpu_state_dict()
    size = spec["tile"] or spec["size"]
    _, noisy = O.synthetic_xray(1, size, size, seed=7)
    t0 = time.perf_counter()
    with torch.no_grad():
        t_un = []
        x = noisy.clone()
        for i in range(unet_evals):
            t1 = time.perf_counter()
            O.unet_forward(sd, x, noisy, torch.full((1,), 49 - i, dtype=torch.long), prefix="diffusion_unet.")
            t_un.append(time.perf_counter() - t1)
        t1 = time.perf_counter(); naf = O._sanitize(O.nafnet_forward(sd, noisy, prefix="nafnet.")); t_naf = time.perf_counter() - t1
        t1 = time.perf_counter(); mask = O._sanitize(O.router_forward(sd, noisy, "router.")); t_r = time.perf_counter() - t1
        t1 = time.perf_counter(); O.fusion_forward(sd, naf, noisy, mask, "fusion."); t_f = time.perf_counter() - t1
    spent = time.perf_counter() - t0
    per_img = spec["tiles"] * (spec["evals"] * sorted(t_un)[len(t_un) // 2] + t_naf + t_r + t_f)
    out = {"ips": 1.0 / per_img, "spent_s": spent, "t_unet_median_s": sorted(t_un)[len(t_un) // 2], "t_rest_s": t_naf + t_r + t_f}
    if full_image:
        t1 = time.perf_counter()
        with torch.no_grad():
            O.hybrid_forward(sd, noisy, 50, 50)
        out["full_image_s"] = time.perf_counter() - t1
        out["full_image_extrapolated_s"] = 50 * out["t_unet_median_s"] + out["t_rest_s"]
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    spec = workload_spec(args, world)
    cores = cpu_threads()
    vals, spent, full = [], [], None
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(spec, CPU_UNET_EVALS, full_image=(args.full_image and i == args.warmup + args.steps - 1))
        if i >= args.warmup:
            vals.append(r["ips"]); spent.append(r["spent_s"])
        if "full_image_s" in r:
            full = {"measured_s": r["full_image_s"], "extrapolated_s": r["full_image_extrapolated_s"],
                    "ratio": r["full_image_extrapolated_s"] / r["full_image_s"]}
    v = sum(vals) / len(vals)
    sample = (f"batch 1 @512x512: {CPU_UNET_EVALS} UNet evals + 1 NAFNet + 1 router + 1 fusion per step (~{sum(spent) / len(spent):.1f} s of CPU work), "
              f"EXTRAPOLATED to one image as {spec['tiles']} x ({spec['evals']} * median(t_unet) + rest)")
    line = {
        "impl": "reference", "metric": spec["metric"], "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "extrapolated": True, "sample_ms_per_step": 1000.0 * sum(spent) / len(spent),
        "config": {"workload": spec["label"],
                   "note": "CPU port of the reference algorithm (oracle/xrd_oracle.py: the same ATen ops the reference classes dispatch to); the "
                           "reference is a Python program, nothing to compile.  value and ms_per_step are per-image figures EXTRAPOLATED from the "
                           "bounded sample each step runs (sample_ms_per_step is the measured wall time of one step); no product code is imported"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if full:
        line["full_image_check"] = full
    try:        # self-check: this arm must not have mapped the product's native library
        line["native_so_loaded"] = sorted({ln.split()[-1] for ln in open("/proc/self/maps") if "libxrd" in ln})
    except Exception:
        pass
    emit(line)


# ------------------------------------------------------------------------------------------------
# roofline of the dominant kernel class: every conv layer of one UNet evaluation, through the kernel the engine dispatches
# ------------------------------------------------------------------------------------------------
def unet_conv_layers(s):
    """(count per eval, Cin, Cout, H, k, stride, flags): the convolutions of UNetDiffusion.forward at the reference topology.
    flags: "gn" = the engine applies GroupNorm+SiLU of the input inside the conv kernel (engine.cu resblock); "cat" = the input is
    the virtual concat of two tensors (torch.cat([x, skip]) of the up path, never materialised)."""
    L = [(5, 48, 48, s, 3, 1, "gn"), (1, 96, 48, s, 3, 1, "gn cat"), (2, 48, 48, s // 2, 3, 1, "gn"), (1, 48, 96, s // 2, 3, 1, "gn"),
         (4, 96, 96, s // 2, 3, 1, "gn"), (1, 192, 96, s // 2, 3, 1, ""), (1, 96, 96, s // 2, 3, 1, ""), (1, 96, 48, s // 2, 3, 1, "gn cat"),
         (1, 192, 48, s // 2, 3, 1, ""), (2, 96, 96, s // 4, 3, 1, "gn"), (1, 96, 144, s // 4, 3, 1, ""), (4, 144, 144, s // 4, 3, 1, ""),
         (1, 288, 144, s // 4, 3, 1, ""), (1, 144, 144, s // 4, 3, 1, ""), (1, 192, 96, s // 4, 3, 1, ""), (1, 288, 96, s // 4, 3, 1, ""),
         (2, 144, 144, s // 8, 3, 1, ""), (1, 144, 192, s // 8, 3, 1, ""), (10, 192, 192, s // 8, 3, 1, ""), (3, 384, 192, s // 8, 3, 1, ""),
         (1, 192, 192, s // 8, 3, 1, ""), (1, 384, 144, s // 8, 3, 1, ""), (1, 288, 144, s // 8, 3, 1, ""),
         (6, 192, 576, s // 8, 1, 1, ""), (6, 192, 192, s // 8, 1, 1, ""),
         (1, 48, 48, s, 3, 2, ""), (1, 96, 96, s // 2, 3, 2, ""), (1, 144, 144, s // 4, 3, 2, ""),
         # the 15 res_conv 1x1s (every ResidualBlock whose channel count changes)
         (1, 48, 96, s // 2, 1, 1, ""), (1, 96, 144, s // 4, 1, 1, ""), (1, 144, 192, s // 8, 1, 1, ""), (3, 384, 192, s // 8, 1, 1, ""),
         (1, 384, 144, s // 8, 1, 1, ""), (1, 288, 144, s // 8, 1, 1, ""), (1, 288, 144, s // 4, 1, 1, ""), (1, 288, 96, s // 4, 1, 1, ""),
         (1, 192, 96, s // 4, 1, 1, ""), (1, 192, 96, s // 2, 1, 1, ""), (1, 192, 48, s // 2, 1, 1, ""), (1, 96, 48, s // 2, 1, 1, ""),
         (1, 96, 48, s, 1, 1, "")]
    return L


def conv_roofline(mode, dev, batch, size):
    """Time every distinct conv layer shape of one UNet evaluation live (CUDA events on the launching stream, via the C ABI op
    hook that launches the very kernel the engine dispatches, including the fused-GroupNorm variants) and return
    launch-weighted achieved TFLOP/s = sum(algorithmic FLOPs) / sum(kernel time)."""
    import ctypes as C
    from xrd_b200 import _lib
    lib = _lib.load()
    cfg = _lib.default_config()
    h = C.c_void_p()
    _lib.check(lib.xrd_create(dev.index or 0, C.byref(cfg), C.byref(h)))
    _lib.check(lib.xrd_set_mode(h, {"bf16": 0, "fp32": 1, "fp16": 2}[mode]))
    tot_ms, tot_fl, launches = 0.0, 0.0, 0
    per_kernel = {}
    ms = C.c_float()
    for cnt, ci, co, hh, k, st, flags in unet_conv_layers(size):
        x = torch.randn(batch, ci, hh, hh, device=dev)
        w = torch.randn(co, ci, k, k, device=dev) * 0.05
        b = torch.zeros(co, device=dev)
        ho = (hh + 2 * (k // 2) - k) // st + 1
        y = torch.empty(batch, co, ho, ho, device=dev)
        # same dispatch as the engine (engine.cu conv() / resblock()): the N-stacked row-ring kernel conv3s for 3x3 / stride 1 layers
        # with 48 or 96 outputs and up to 96 inputs on maps >= 128 wide (GroupNorm+SiLU of the input applied inside it in the residual
        # blocks, the up path's concat read from its two sources), conv3 (other W % 128 == 0 layers), conv3w (W == 64), conv1 (1x1),
        # else the per-tap kernel (the three stride-2 convolutions)
        if k == 3 and st == 1 and hh % 128 == 0 and hh >= 128 and co in (48, 96) and ci in (48, 96):
            gn, cat = "gn" in flags, "cat" in flags
            impl = (18 if gn else 17) if cat else (16 if gn else 15)
            kern = "k_conv3s" + ("+gn" if gn else "") + ("(cat)" if cat else "")
        elif k == 3 and st == 1 and hh % 128 == 0 and co in (48, 96, 144):
            impl, kern = 2, "k_conv3"
        elif k == 3 and st == 1 and hh == 64 and co in (144, 192):
            impl, kern = 7, "k_conv3w"
        elif k == 1 and st == 1:
            impl, kern = 5, "k_conv1"
        else:
            impl, kern = 1, "k_conv_tc"
        _lib.check(lib.xrd_op_conv2d(h, impl, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                     C.c_void_p(y.data_ptr()), batch, ci, hh, hh, co, k, st, k // 2, None))
        _lib.check(lib.xrd_op_time_last(h, 5, C.byref(ms), None))
        tot_ms += cnt * ms.value
        fl = cnt * 2.0 * batch * ho * ho * co * ci * k * k
        tot_fl += fl
        launches += cnt
        per = per_kernel.setdefault(kern, [0.0, 0.0, 0])
        per[0] += fl; per[1] += cnt * ms.value; per[2] += cnt
        del x, w, y
    lib.xrd_destroy(h)
    torch.cuda.empty_cache()
    by_kernel = {k: {"tflops": v[0] / v[1] / 1e9, "ms_per_eval": v[1], "launches_per_eval": v[2]} for k, v in per_kernel.items()}
    return tot_fl / tot_ms / 1e9, tot_ms, launches, tot_fl, by_kernel


def insitu_profile(model, x_dev, evals, conv_flops_per_eval):
    """Per-kernel times of the sampler measured WHERE THE KERNELS RUN: one whole reverse loop (all `evals` UNet evaluations of
    the step, the launch sequence the CUDA graph replays) launched eagerly with every kernel bracketed by two CUDA events on its
    stream (include/xrd.h: xrd_profile_begin / _end).  A kernel timed in a loop of its own runs at the clock the power cap
    allows THAT kernel (conv3s 48->48 @512x512 alone: 1.24 GHz at 990 W); here it runs between its real neighbours at the clock of
    the real mix, which is what the step is made of.  Returns the conv class's launch-weighted TFLOP/s and the table."""
    from xrd_b200 import _lib
    w = model.diffusion_wrapper
    model.use_cuda_graph = False
    model.diffusion_unet.use_cuda_graph = False
    steps = model.inference_diffusion_steps
    try:
        w.denoise(x_dev, steps)                  # eager warm-up (plans, packs)
        torch.cuda.synchronize()
        _lib.profile_begin()
        try:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            w.denoise(x_dev, steps)
            e1.record()
            torch.cuda.synchronize()
        finally:
            prof = _lib.profile_end()
        loop_ms = e0.elapsed_time(e1)
    finally:
        model.use_cuda_graph = True
        model.diffusion_unet.use_cuda_graph = True
    conv = ("k_conv3s", "k_conv3", "k_conv3r", "k_conv3w", "k_conv1", "k_conv_tc")
    tot = sum(v[1] for v in prof.values())
    conv_ms = sum(v[1] for k, v in prof.items() if k in conv)
    table = {k: {"launches_per_eval": v[0] / evals, "ms_per_eval": v[1] / evals, "share": v[1] / tot} for k, v in prof.items() if v[1] / tot >= 0.002}
    return {"how": "eager replay of the step's reverse loop, two CUDA events around every launch (xrd_profile_begin/_end), same batch and "
                   "inputs as the timed region, taken right after it; the conv class is the tcgen05 kernels of unet_conv_layers() "
                   "(the two-plane first conv, k_first_conv_mma, and out_conv are HBM-bound edge kernels: neither their FLOPs nor their "
                   "time are in it)",
            "conv_tflops": conv_flops_per_eval * evals / conv_ms / 1e9, "conv_ms_per_eval": conv_ms / evals,
            "kernels_ms_per_eval": tot / evals, "eager_loop_ms": loop_ms, "evals": evals, "by_kernel": table}


_REAL_STDOUT = None


def _claim_stdout():
    """This program prints ONE JSON line on stdout.  Native libraries write there too (NCCL's version banner at WARN level and
    above): send file descriptor 1 to stderr for the whole run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="hybrid512", choices=["hybrid512", "tiled1024"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default 16; 8 for tiled1024)")
    ap.add_argument("--global-batch", type=int, default=0, help="fixed batch sharded by image over all ranks (strong scaling, configs[3])")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--mode", default=os.environ.get("XRD_MODE", "fp16"), choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-layer conv timing (secondary workloads)")
    ap.add_argument("--full-image", action="store_true", help="reference arm: also time one whole un-extrapolated DDIM-50 image (~30-120 s)")
    args = ap.parse_args()
    if not args.batch:
        args.batch = 8 if args.workload == "tiled1024" else 16

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import xrd_b200
    import synthetic_data as SY
    spec = workload_spec(args, world)
    model, run = build_model(dev, args.mode, spec)
    S = spec["size"]
    if args.global_batch:                      # strong scaling: this rank's slice of ONE fixed batch
        lo, hi = xrd_b200.shard_bounds(args.global_batch, rank, world)
        B, total_images = hi - lo, args.global_batch
        _, noisy = SY.synthetic_xray(max(B, 1), S, S, seed=7 + rank)
        noisy = noisy[:B]
    else:                                      # weak scaling: a fixed batch per GPU
        B, total_images = args.batch, args.batch * world
        _, noisy = SY.synthetic_xray(B, S, S, seed=7 + rank)
    bmax = B
    if world > 1 and args.global_batch:
        bmax = max(xrd_b200.shard_bounds(args.global_batch, r, world)[1] - xrd_b200.shard_bounds(args.global_batch, r, world)[0] for r in range(world))
    x_dev = noisy.to(dev)
    x_pin = noisy.pin_memory()
    out_pin = torch.empty_like(noisy).pin_memory()
    pad = torch.zeros((bmax, 1, S, S), device=dev) if world > 1 else None
    gathered = [torch.empty_like(pad) for _ in range(world)] if world > 1 else None

    def gather(y):
        pad[:B] = y                                # ragged slices: every rank contributes a buffer of the largest slice
        dist.all_gather(gathered, pad)             # output gather only; nothing inside the sampler communicates

    def step_resident():
        y = run(x_dev)
        if world > 1:
            gather(y)
        return y

    def step_e2e():
        xd = x_pin.to(dev, non_blocking=True)
        y = run(xd)
        if world > 1:
            gather(y)
        out_pin.copy_(y, non_blocking=True)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        step_resident()
    sync_all()
    launches0 = xrd_b200.native_kernel_launches()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = xrd_b200.native_kernel_launches() - launches0
    # ---- e2e: pinned host -> device -> model -> host, every step ----
    for _ in range(2):
        step_e2e()
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    ips = total_images * args.steps / (ms / 1000.0)
    ips_e2e = total_images * args.steps / (ms_e2e / 1000.0)
    do_roof = args.mode != "fp32" and not args.no_roofline
    conv_tf, conv_ms, conv_launches, conv_fl, conv_by = conv_roofline(args.mode, dev, 16, spec["tile"] or S) if do_roof else (0.0, 0.0, 0, 0.0, {})
    step_tf = ips / world * spec["gf_per_image"] / 1000.0
    traffic = top_kernel_traffic(16, spec["tile"] or S)
    insitu = None
    if do_roof and not spec["tile"] and B == 16:
        try:
            insitu = insitu_profile(model, x_dev, spec["evals"], conv_fl)
        except Exception as e:  # noqa: BLE001  -- the measurement aid must not cost the bench line
            insitu = {"error": str(e)[:300]}
    line = {
        "metric": spec["metric"], "value": ips, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.global_batch else "weak",
        "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.mode], "data": "synthetic",
        "config": {"workload": spec["label"], "global_batch": total_images, "images_this_rank": B, "unet_evals_per_image": spec["evals"] * spec["tiles"],
                   "accumulate": "fp32",
                   "l2": "working set per step is several GB (>> 126 MB L2): inputs larger than L2, no explicit flush",
                   "parallelism": f"image-sharded x{world}, output all_gather only"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": B * S * S * 4, "d2h_bytes_per_step": B * S * S * 4},
        "roofline": {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM convolutions (k_conv3s, k_conv3, k_conv3w, k_conv1, k_conv_tc): every conv layer shape of one UNet "
                               "evaluation at batch 16 / 512x512, each timed live with CUDA events through the C-ABI op hook on the kernel variant the engine "
                               "dispatches (fused GroupNorm+SiLU where the network fuses it), launch weighted",
                     "achieved": conv_tf if do_roof else None, "peak": pk["burst"], "unit": "TFLOP/s",
                     "frac": (conv_tf / pk["burst"]) if (do_roof and pk["burst"]) else None,
                     "peak_kind": f"bf16 dense burst, {pk['src']}",
                     "traffic": traffic["dram_bytes"] if traffic else None, "traffic_source": traffic if traffic else "no ncu capture of this tree for this shape",
                     "by_kernel": conv_by, "flops_per_eval": conv_fl, "ms_per_eval_isolated": conv_ms, "launches_per_eval": conv_launches,
                     "in_situ": insitu},
        "roofline_step": {"bound": "tensor", "achieved": step_tf, "peak": pk["sustained"], "unit": "TFLOP/s",
                          "frac": step_tf / pk["sustained"], "note": f"images/s x {spec['gf_per_image']:.0f} GF algorithmic per image, of sustained measured peak"},
    }
    if not args.no_cpu_baseline and world == 1:          # rank 0 at N=1 only (the N>1 runs keep their wall time for the GPUs)
        cores = cpu_threads()
        r = cpu_reference_sample(spec, CPU_UNET_EVALS)
        line["cpu_baseline"] = {"value": r["ips"], "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"batch 1 @512x512: {CPU_UNET_EVALS} UNet evals + NAFNet + router + fusion ({r['spent_s']:.1f} s of CPU work), "
                                          f"EXTRAPOLATED as {spec['tiles']} x ({spec['evals']} * median(t_unet) + rest)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
