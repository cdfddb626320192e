#!/usr/bin/env python3
"""Text summary of an `ncu -i X.ncu-rep --page raw --csv` export: the metrics DESIGN.md quotes, per captured launch.
python tools/ncu_summary.py raw.csv > profiles/rNN_..._ncu.txt"""
import csv, sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "lts__t_sector_hit_rate.pct",
]


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, units = rows[head], rows[head + 1]
    kn = h.index("Kernel Name")
    for r in rows[head + 2:]:
        if len(r) <= kn:
            continue
        print(f"## {r[kn][:110]}")
        for m in METRICS:
            if m in h:
                i = h.index(m)
                print(f"  {m:<86s} {r[i]} {units[i]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
