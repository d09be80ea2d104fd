"""Per-kernel / per-grid breakdown of the LAST UNet evaluation in an ncu launch list (gpu__time_duration csv)."""
import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches_ddim.csv"
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if "k_conv_smallcin" in r["Kernel Name"] or "k_im2col_3x3_2ch" in r["Kernel Name"] or "k_first_conv_mma" in r["Kernel Name"]]   # first kernel of an evaluation
sel = rows[idx[-1]:]
agg = collections.OrderedDict(); tot = 0
for r in sel:
    name = re.sub(r"[<(].*", "", r["Kernel Name"]).replace("void ", "").replace("xrd::", "").replace("(anonymous namespace)::", "")
    t = float(r["Metric Value"].replace(",", "")) / 1e3
    tot += t
    key = (name, r["Grid Size"])
    agg.setdefault(key, [0, 0.0]); agg[key][0] += 1; agg[key][1] += t
print(f"last evaluation: {tot/1e3:.2f} ms over {len(sel)} launches")
byk = collections.Counter()
for (n, g), v in agg.items(): byk[n] += v[1]
for n, v in byk.most_common(): print(f"  {n:24s} {v/1e3:7.2f} ms {100*v/tot:5.1f}%")
print()
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    print(f"{k[0]:22s} grid={k[1]:16s} n={v[0]:3d} total={v[1]:8.1f} us avg={v[1]/v[0]:7.1f}")
