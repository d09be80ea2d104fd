"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/ncu_summary.py launches.csv [skip_first_n]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for i, row in enumerate(csv.DictReader(lines)):
        if i < skip:
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u in ("s", "second") else v
        name = re.sub(r"[<(].*", "", row["Kernel Name"]).replace("void ", "")
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T/1e3:.2f} ms over {sum(cnt.values())} launches (skipped first {skip})")
    print(f"{'kernel':28s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in tot.most_common(30):
        print(f"{k:28s} {cnt[k]:8d} {v/1e3:10.2f} {100*v/T:6.2f}% {v/cnt[k]:9.1f}")


if __name__ == "__main__":
    main()
