"""profiles/top_kernel_traffic.json (read by bench.py's `roofline.traffic`) from the compact CSV of an `ncu --set full` capture
(tools/ncu_export.py): dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel at the benched shape.
    python tools/ncu_traffic.py hot.csv 'k_conv3s<__half, 3, 0, 8, 1, 1>' 16 512 out.json"""
import csv, json, sys

src, pat, batch, size, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(open(src)))
head, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[2:]:
    if pat in r[0]:
        rd = float(r[head.index("dram__bytes_read.sum")]) * scale[units[head.index("dram__bytes_read.sum")]]
        wr = float(r[head.index("dram__bytes_write.sum")]) * scale[units[head.index("dram__bytes_write.sum")]]
        d = {"kernel": r[0], "dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "batch": batch, "size": size,
             "duration_us_under_ncu": float(r[head.index("gpu__time_duration.sum")]),
             "capture": "ncu --set full --clock-control none --import-source on, one launch at the batch-16 / 512x512 shape (tools/ncu_targets_r2.py); "
                        "table: profiles/r02_ncu_hot_kernels.csv"}
        json.dump(d, open(out, "w"), indent=1)
        print(d)
        break
else:
    sys.exit(f"no kernel matching {pat!r} in {src}")
