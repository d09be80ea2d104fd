"""Builds libxrd.so in-tree with plain nvcc for sm_100a (no torch extension machinery:
the library has a C ABI and links only the static CUDA runtime)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libxrd.so")
# Test-only twin of the library: conv3s.cu compiled with XRD_C3S_RACE_TEST (its TMA producer fetches rows out of order, microseconds
# apart) so that the GPU suite can run the row protocol under an exaggerated out-of-order completion of TMA loads
# (tests/test_gpu_parity.py::test_conv3s_row_protocol_under_fault_injection).  Never loaded by the package.
OUT_FAULTINJ = os.path.join(_HERE, "libxrd_faultinj.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--use_fast_math_off_placeholder",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(OUT) or not os.path.exists(OUT_FAULTINJ):
        return True
    t = min(os.path.getmtime(OUT), os.path.getmtime(OUT_FAULTINJ))
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(_ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libxrd.so (one object per source, in parallel)."""
    if not force and not _stale():
        return OUT
    nvcc = _nvcc()
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if "placeholder" not in f]
    common = flags + ["-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", os.path.join(_ROOT, "include"), "-I", CSRC]
    if verbose:
        common += ["-Xptxas", "-v"]
    if os.environ.get("XRD_TRACE"):          # in-kernel clock64 traces and experiment switches of conv3 / conv3r (tools/c3_trace.py, c3r_dbg.py)
        common += ["-DXRD_TRACE"]
    procs = []
    objs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, "-c", src, "-o", obj] + common,
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    fi_obj = os.path.join(objdir, "conv3s_faultinj.o")
    procs.append((os.path.join(CSRC, "conv3s.cu") + " [fault injection]",
                  subprocess.Popen([nvcc, "-c", os.path.join(CSRC, "conv3s.cu"), "-o", fi_obj, "-DXRD_C3S_RACE_TEST"] + common,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0):
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, "-shared", "-o", OUT] + objs + flags + ["-cudart", "static", "-Xlinker", "--no-undefined",
                                                           "-lpthread", "-ldl", "-lrt"]
    link_fi = [nvcc, "-shared", "-o", OUT_FAULTINJ] + [o for o in objs if os.path.basename(o) != "conv3s.o"] + [fi_obj] + flags + \
        ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
    for cmd in (link, link_fi):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
