#!/usr/bin/env python
"""Dynamic (executed) instruction mix of one kernel of an ncu report taken with --import-source on.
Usage: tools/ncu_dynmix.py <report.ncu-rep> <kernel-name-substring>"""
import collections
import csv
import subprocess
import sys
rep = sys.argv[1]; kern = sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-kernel-base","function"],capture_output=True,text=True).stdout
for b in out.split('"Kernel Name",')[1:]:
    lines = b.splitlines(); name = lines[0].strip('",')
    if kern not in name: continue
    rows = list(csv.reader(lines[1:])); hdr = rows[0]; data = [r for r in rows[1:] if len(r)==len(hdr)]
    ix = {h:i for i,h in enumerate(hdr)}
    mix = collections.Counter(); tot = 0
    for r in data:
        src = r[ix["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS','STS','LDG','STG','BAR')) and '.' in op else '')
        n = int(r[ix["Instructions Executed"]]); mix[op] += n; tot += n
    print("==", name, "warp-instr", tot)
    for op, n in mix.most_common(20):
        print(f"  {op:14s} {100*n/tot:5.1f}%  {n}")
    break
