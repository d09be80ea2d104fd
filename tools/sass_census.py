"""SASS mnemonic census of the shipped objects (no GPU needed): which kernels files really carry tcgen05 MMAs (UTCHMMA), TMA tensor /
bulk copies (UTMALDG / UBLKCP), TMEM loads and stores (LDTM / STTM), tcgen05.commit (UTCBAR), mbarrier traffic (SYNCS) and warp-level
MMAs (HMMA).   python tools/sass_census.py > profiles/r02_sass_census.txt"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "medical-image-denoising-using-diffusion_b200", "build")
MN = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2"]
print(f"{'object':22s} {'kernels':>7s} " + " ".join(f"{m:>8s}" for m in MN))
for o in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    nk = len(re.findall(r"^\s*Function :", sass, flags=re.M))
    if not nk:
        continue
    cnt = [len(re.findall(r"(?<![A-Z])" + re.escape(m) + r"\b", sass)) for m in MN]
    print(f"{os.path.basename(o):22s} {nk:7d} " + " ".join(f"{c:8d}" for c in cnt))
