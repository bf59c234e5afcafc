#!/usr/bin/env python
"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump to the per-launch metrics the roofline statements rest on
(B200_PROFILING.md): duration, DRAM bytes read / written, DRAM throughput %, tensor-pipe active %, issue-slot use, occupancy,
registers.    python tools/ncu_summary.py gpurun_out/r02_hbm_full_raw.csv > profiles/r02_hbm_ncu_summary.csv"""
import csv
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([n + (f" [{units[i]}]" if units[i] else "") for i, n in cols])
    for r in data:
        if len(r) < len(hdr):
            continue
        w.writerow([r[i][:90] for i, _ in cols])


if __name__ == "__main__":
    main()
