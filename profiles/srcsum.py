#!/usr/bin/env python
"""Summarise `ncu --page source --csv` for one kernel: stall-reason totals and the hottest SASS lines.
   python profiles/srcsum.py REPORT.ncu-rep KERNEL_REGEX [TOP]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + rx], capture_output=True, text=True).stdout
lines = out.splitlines()
# several kernels may match: split on the "Kernel Name" rows, keep the first
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
seg = lines[starts[0] + 1:(starts[1] if len(starts) > 1 else len(lines))]
print(lines[starts[0]][:160])
rows = list(csv.reader(seg))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); samples = 0
body = rows[1:]
for r in body:
    s = int(r[ix['# Samples']] or 0); samples += s
    for h in stall:
        tot[h] += int(r[ix[h]] or 0)
print('samples', samples, 'instructions', len(body), 'warp-insts executed', sum(int(r[ix['Instructions Executed']] or 0) for r in body))
print('stalls:', ', '.join(f"{h[6:]} {100*v/max(1,samples):.1f}%" for h, v in tot.most_common(8)))
for r in sorted(body, key=lambda r: -int(r[ix['# Samples']] or 0))[:top]:
    s = int(r[ix['# Samples']] or 0)
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall), reverse=True)[:2]
    print(f"{100*s/max(1,samples):5.1f}%  {r[ix['Source']].strip()[:70]:70s} exec={r[ix['Instructions Executed']]:>8s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
