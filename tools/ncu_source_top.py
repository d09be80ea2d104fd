"""Top stall sites of an `ncu --page source --csv` export:  python tools/ncu_source_top.py one_source.csv [n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    ix = {k: i for i, k in enumerate(h)}
    body = [r for r in rows[hi + 1:] if len(r) == len(h)]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = sum(f(r, "# Samples") for r in body)
    stall = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    print(rows[0][1][:100] if rows[0] else "", "| samples", tot, "| instructions", len(body))
    agg = sorted(((sum(f(r, k) for r in body), k) for k in stall), reverse=True)[:8]
    print("  ".join(f"{k[6:]}={100 * v / tot:.1f}%" for v, k in agg))
    for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:n]:
        st = max(((f(r, k), k) for k in stall))
        print(f"{r[0][-5:]} {100 * f(r, '# Samples') / tot:5.1f}%  exec={f(r, 'Instructions Executed'):10.0f}  {r[1][:72]:72s} {st[1][6:]}")


if __name__ == "__main__":
    main()
