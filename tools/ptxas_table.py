"""Per-kernel resource table from `nvcc -Xptxas -v` logs (registers, spills, static shared memory), kernel names demangled.
    for f in conv3s conv3 ...; do nvcc -c csrc/$f.cu ... -Xptxas -v > /tmp/px_$f.log 2>&1; done
    python tools/ptxas_table.py /tmp/px_*.log > profiles/r02_ptxas_resources.txt"""
import re
import subprocess
import sys

rows = []
for path in sys.argv[1:]:
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r".*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt, flags=re.S):
        rows.append((m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(6) or 0)))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"{'regs':>4s} {'stack':>5s} {'spill st':>8s} {'spill ld':>8s} {'smem':>6s}  kernel")
for (_, regs, stack, st, ld, smem), n in zip(rows, names):
    n = re.sub(r"^void xrd::|\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(.*$", "", n)
    print(f"{regs:4d} {stack:5d} {st:8d} {ld:8d} {smem:6d}  {n}")
