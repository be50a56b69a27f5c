#!/usr/bin/env python
"""stdin: `ncu -i rep --page raw --csv`; stdout: one row per captured launch with the metrics the roofline uses."""
import csv
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "lts__t_bytes.sum", "sm__inst_executed_pipe_tensor.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
if len(rows) < 3:
    sys.exit("no launches captured")
hdr = rows[0]
cols = [(k, hdr.index(k)) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
w.writerow(["%s [%s]" % (k, rows[1][i]) if rows[1][i] else k for k, i in cols])   # rows[1] holds the units
for r in rows[2:]:
    w.writerow([r[i] for _, i in cols])
