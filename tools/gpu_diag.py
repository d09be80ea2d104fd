"""Run every GPU check in its own subprocess (a faulting kernel must not poison the rest) and
write gpurun_out/diag.json.   python tools/gpu_diag.py [name ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def main():
    from gpu_checks import CHECKS
    names = sys.argv[1:] or list(CHECKS)
    results = {}
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_checks.py"), name], capture_output=True,
                               text=True, timeout=float(os.environ.get("XRD_CHECK_TIMEOUT", "240")), cwd=ROOT)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if r.returncode == 0 and line:
                results.update(json.loads(line[-1][7:]))
            else:
                results[name] = {"ERROR": (r.stderr or r.stdout)[-1500:], "rc": r.returncode}
        except subprocess.TimeoutExpired:
            results[name] = {"ERROR": "timeout"}
        dt = time.time() - t0
        v = results[name]
        if "ERROR" in v:
            print(f"[{name}] FAILED ({dt:.1f}s): {v['ERROR'][-600:]}", flush=True)
        else:
            print(f"[{name}] ({dt:.1f}s) " + " ".join(f"{k}={val:.3g}" if isinstance(val, float) else f"{k}={val}" for k, val in v.items()),
                  flush=True)
        with open(os.path.join(OUT, "diag.json"), "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
