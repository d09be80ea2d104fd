"""Compact per-kernel CSV from an `ncu --set full` report (the columns profiles/README.md quotes).
    python tools/ncu_export.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = [head.index(k) for k in KEEP if k in head]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            w.writerow([r[i] for i in idx])
    print(f"{len(body)} kernels x {len(idx)} columns -> {out}")


if __name__ == "__main__":
    main()
