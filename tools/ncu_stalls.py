#!/usr/bin/env python3
"""Top stall sites of each kernel in an .ncu-rep (source page, SASS): `ncu_stalls.py REP [topN]`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern, hdr, body = None, None, []
def flush():
    if not body:
        return
    si = hdr.index('# Samples'); src = hdr.index('Source'); ex = hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[si] or 0) for r in body)
    print(f'== {kern[:110]}  total samples {tot}, instructions {len(body)}')
    agg = {}
    for r in body:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print('   by reason:', ', '.join(f'{k[6:]}={v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ranked = sorted(enumerate(body), key=lambda ir: -int(ir[1][si] or 0))[:top]
    for idx, r in sorted(ranked):
        reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f'   #{idx:5d} {int(r[si]):7d} ({100*int(r[si])/max(tot,1):5.1f}%) exec {r[ex]:>8s}  {r[src].strip()[:70]:70s} {reasons}')
for r in rows:
    if r and r[0] == 'Kernel Name':
        flush(); kern = r[1]; hdr = None; body = []
    elif r and r[0] == 'Address':
        hdr = r
    elif hdr and len(r) == len(hdr):
        body.append(r)
flush()
