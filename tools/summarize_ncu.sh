#!/bin/bash
# Summarise a .ncu-rep (run here, no GPU): key metrics per profiled launch -> stdout
REP=$1
ncu -i "$REP" --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]
idx = {h: i for i, h in enumerate(hdr)}
for d in data:
    print("-" * 100)
    for w in want:
        if w in idx:
            print(f"{w:70s} {d[idx[w]]:>20s} {units[idx[w]]}")
'
