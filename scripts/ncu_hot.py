#!/usr/bin/env python
"""Top SASS instructions of an .ncu-rep by stall samples, with executed counts and average active threads.
usage: python scripts/ncu_hot.py file.ncu-rep [N]"""
import csv, io, subprocess, sys
path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {k: hdr.index(k) for k in ('Address', 'Source', '# Samples', 'Instructions Executed', 'Avg. Threads Executed', 'Warp Stall Sampling (All Samples)')}
body = [r for r in rows[2:] if len(r) > ix['# Samples'] and r[ix['# Samples']].isdigit()]
tot_s = sum(int(r[ix['# Samples']]) for r in body); tot_i = sum(int(r[ix['Instructions Executed']]) for r in body)
print(f'{len(body)} SASS lines, {tot_s} samples, {tot_i} warp instructions')
# opcode histogram by executed instructions
hist = {}
for r in body:
    op = r[ix['Source']].split()[0] if r[ix['Source']].split() else '?'
    if op.startswith('@'):
        op = r[ix['Source']].split()[1]
    op = op.split('.')[0]
    h = hist.setdefault(op, [0, 0, 0.0])
    h[0] += int(r[ix['Instructions Executed']]); h[1] += int(r[ix['# Samples']]); h[2] += int(r[ix['Instructions Executed']]) * float(r[ix['Avg. Threads Executed']] or 0)
print('opcode: %warp-instr  %samples  avg-threads')
for op, (i, sm, th) in sorted(hist.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f'  {op:10s} {100*i/tot_i:5.1f}%  {100*sm/tot_s:5.1f}%  {th/max(i,1):5.1f}')
print('hottest lines by samples:')
for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][ix['# Samples']]))[:n]:
    print(f"  #{idx:4d} {int(r[ix['# Samples']]):6d} smp  {int(r[ix['Instructions Executed']]):10d} exe  {r[ix['Avg. Threads Executed']]:>5s} thr  {r[ix['Source']].strip()[:70]}")
