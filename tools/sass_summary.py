#!/usr/bin/env python
"""Instruction mix (SASS mnemonics) of the hot kernels, from the objects the library is linked from:
    python tools/sass_summary.py > profiles/r2_sass_summary.txt
Counts are static (per kernel body, not executed counts); the point is WHICH instructions the hand-written kernels
compile to: packed fp32 pairs (FFMA2/FADD2/FMUL2), 128-bit global accesses, cp.async (LDGSTS), bulk L2 prefetch
(UBLKPF), shuffles, and the absence of local-memory traffic (LDL/STL = spills)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "torchfsm_b200", "csrc", "_build")
WANT = [("fsm_kernels_1024.o", r"k_pass_ixIfNS_6FftCfgILi1024.*ELi3EE"), ("fsm_kernels_1024.o", r"k_pass_physzIf.*Li1024"),
        ("fsm_kernels_1024.o", r"k_pass_fxIfNS_6FftCfgILi1024.*ELi1EE"),
        ("fsm_kernels_512.o", r"k_pass_ixIfNS_6FftCfgILi512.*ELi1EE"), ("fsm_kernels_512.o", r"k_pass_midIfNS_6FftCfgILi512.*ELi1ELi0ELi0EE"),
        ("fsm_kernels_512.o", r"k_pass_midIfNS_6FftCfgILi512.*ELi1ELi1ELi0EE"), ("fsm_kernels_512.o", r"k_pass_midIfNS_6FftCfgILi512.*ELin1ELi0ELi1EE"),
        ("fsm_kernels_512.o", r"k_pass_physIfNS_6FftCfgILi512.*ELi1ELi3ELi8EE"), ("fsm_kernels_512.o", r"k_pass_fxIfNS_6FftCfgILi512.*ELi3EE"),
        ("fsm_kernels_256.o", r"k_pass_physzIf.*Li256.*ELi7EE"), ("fsm_plan.o", r"k_spectral_mapIf"), ("fsm_plan.o", r"k_combine_onlyIf")]
KEYS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "LDS", "STS", "LDGSTS",
        "UBLKPF", "SHFL", "BAR", "LDL", "STL", "IMAD", "MUFU", "UTMALDG", "UTMASTG"]
cache = {}
for obj, pat in WANT:
    path = os.path.join(BUILD, obj)
    if path not in cache:
        cache[path] = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    txt = cache[path]
    funcs = re.split(r"\n\s*Function : ", txt)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if not re.search(pat, name):
            continue
        ops = collections.Counter()
        n = 0
        for line in f.split("\n"):
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            n += 1
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k.count(".") and op.startswith(k)):
                    ops[k] += 1
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        demangled = re.sub(r"\(fsm::Geom.*", "", demangled).replace("fsm::", "")
        print(f"{demangled}\n    {n} instructions; " + ", ".join(f"{k} {ops[k]}" for k in KEYS if ops[k] or k in ("LDL", "STL", "SHFL", "UTMALDG")))
        break
    else:
        print(f"[no match for {pat} in {obj}]", file=sys.stderr)
