#!/usr/bin/env python
"""One line per kernel from an .ncu-rep (raw page): time, DRAM bytes, key utilisation, launch config.
   python tools/rawsum.py REPORT.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
K = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
     ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
     ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
     ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1%'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
     ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
     ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
     ('launch__occupancy_limit_registers', 'limR'), ('launch__occupancy_limit_shared_mem', 'limS')]
ix = {k: hdr.index(k) for k, _ in K if k in hdr}
print(f"{'kernel':42s}" + ''.join(f"{n:>9s}" for k, n in K))
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')][:42]
    print(f"{name:42s}" + ''.join(f"{float(r[ix[k]].replace(',', '')):9.1f}" if k in ix else '        -' for k, _ in K))
