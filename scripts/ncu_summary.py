#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu -i ... --page raw --csv) into the few columns profiles/ keeps per kernel launch."""
import csv
import subprocess
import sys

COLS = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'launch__grid_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'sm__cycles_elapsed.max', 'sm__warps_active.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [c for c in COLS if c in idx]
w = csv.writer(sys.stdout)
w.writerow(cols)
w.writerow([units[idx[c]] for c in cols])
for r in rows[2:]:
    w.writerow([r[idx[c]][:70] for c in cols])
