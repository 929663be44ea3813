#!/usr/bin/env python
"""Summarise `ncu --page source --csv` for one kernel launch: stall-reason totals and the hottest SASS lines.
   python tools/srcsum.py REPORT.ncu-rep LAUNCH_INDEX [TOP]"""
import csv, subprocess, sys, collections
rep, idx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', idx, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
print(lines[starts[0]][:160])
rows = list(csv.reader(lines[starts[0] + 1:]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.lower().startswith('stall_') and 'not' not in h.lower()]
if not stall:
    print('columns:', hdr)
body = [r for r in rows[1:] if len(r) == len(hdr) and r[ix['# Samples']].isdigit()]
tot = collections.Counter(); samples = 0
for r in body:
    s = int(r[ix['# Samples']] or 0); samples += s
    for h in stall:
        tot[h] += int(r[ix[h]] or 0)
print('samples', samples, 'instructions', len(body), 'warp-insts executed', sum(int(r[ix['Instructions Executed']] or 0) for r in body))
print('stalls:', ', '.join(f"{h[6:]} {100*v/max(1,samples):.1f}%" for h, v in tot.most_common(10)))
for k, r in enumerate(body):
    r.append(k)
for r in sorted(body, key=lambda r: -int(r[ix['# Samples']] or 0))[:top]:
    s = int(r[ix['# Samples']] or 0)
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall), reverse=True)[:2] if stall else [(0, ''), (0, '')]
    print(f"{100*s/max(1,samples):5.1f}% #{r[-1]:5d} {r[ix['Source']].strip()[:76]:76s} exec={r[ix['Instructions Executed']]:>8s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
