#!/usr/bin/env python
"""Per-CUDA-source-line totals (instructions executed, stall samples) from
   ncu -i REP --page source --print-source sass,cuda --csv > dump.csv"""
import csv, sys, collections
f = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kernel_filter = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(f)))
agg = collections.OrderedDict(); cur_file = None; hdr = None; func = None; tot_i = tot_s = 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; ix = {}; [ix.setdefault(h, i) for i, h in enumerate(hdr)]; continue
    if hdr is None or len(r) != len(hdr): continue
    if kernel_filter and kernel_filter not in (func or ""): continue
    try:
        ins = float(r[ix["Instructions Executed"]] or 0); smp = float(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    key = (cur_file, r[0], r[1].strip()[:110])
    a = agg.setdefault(key, [0.0, 0.0]); a[0] += ins; a[1] += smp
    tot_i += ins; tot_s += smp
print(f"total warp-inst {tot_i:.3g}  samples {tot_s:.0f}")
for (fn, ln, src), (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*ins/tot_i:5.1f}% inst {100*smp/max(tot_s,1):5.1f}% smp  {fn}:{ln:>4}  {src}")
