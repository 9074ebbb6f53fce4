#!/usr/bin/env python
"""Warp-stall sampling per SASS line of one kernel of an ncu report taken with --import-source on.
Usage: tools/ncu_stalls.py <report.ncu-rep> <kernel-name-substring> [top-n]"""
import csv
import subprocess
import sys
rep = sys.argv[1]; kern = sys.argv[2]; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-kernel-base","function"],capture_output=True,text=True).stdout
# split per kernel
blocks = out.split('"Kernel Name",')
for b in blocks[1:]:
    lines = b.splitlines()
    name = lines[0].strip('",')
    if kern not in name: continue
    rows = list(csv.reader(lines[1:]))
    hdr = rows[0]; data = [r for r in rows[1:] if len(r)==len(hdr)]
    ix = {h:i for i,h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("==", name, "samples", tot, "instructions", len(data))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stall_cols}
    print({k[6:]: round(100*v/tot,1) for k,v in sorted(agg.items(), key=lambda kv:-kv[1]) if v*100>tot})
    top = sorted(range(len(data)), key=lambda i:-int(data[i][ix["# Samples"]]))[:topn]
    for i in sorted(top):
        r = data[i]
        st = {h[6:]: int(r[ix[h]]) for h in stall_cols if int(r[ix[h]])>0}
        main = sorted(st.items(), key=lambda kv:-kv[1])[:3]
        print(f"{i:5d} {100*int(r[ix['# Samples']])/tot:5.2f}% {r[ix['Source']].strip()[:60]:60s} {main}")
    break
