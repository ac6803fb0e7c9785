#!/usr/bin/env python
"""SASS of the FIRST kernel launch in an `ncu --page source --print-source sass,cuda --csv` dump, in address order, with
the share of executed warp instructions and the CUDA source line each instruction maps to.
usage: ncu_sass.py dump.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
seen = {}
cur_file = None; cur_line = ('', '')
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if len(r) < 9: continue
    if r[0] != "": cur_line = (r[0], r[1].strip()); continue
    addr = r[2]
    if not addr.startswith("0x") or addr in seen: continue
    try: ins = float(r[7] or 0); smp = float(r[6] or 0)
    except ValueError: continue
    seen[addr] = (int(addr, 16), cur_file, cur_line[0], cur_line[1], r[3], ins, smp)
tot = sum(v[5] for v in seen.values()); tots = sum(v[6] for v in seen.values())
print(f"# {len(seen)} SASS instructions, {tot:.3g} warp-instructions executed, {tots:.0f} samples")
for a, f, ln, src, sass, ins, smp in sorted(seen.values()):
    p = 100 * ins / tot
    if p >= minpct:
        print(f"{p:5.2f}% {100*smp/max(tots,1):5.2f}%s {f}:{ln:>4} | {sass[:70]:70s} | {src[:80]}")
