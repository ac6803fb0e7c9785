#!/usr/bin/env python
"""Per-CUDA-source-line totals of one kernel of an .ncu-rep (captured with --import-source on).
usage: ncu_cuda_lines.py REP kernel_regex [top=40] [launch_skip=0]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "-k", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
cur = None; hdr = None; rows = []
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    ix = {h: i for i, h in reversed(list(enumerate(hdr)))}
    g = lambda k: float(r[ix[k]] or 0) if r[ix[k]] not in ("-", "") else 0.0
    rows.append((cur, int(r[0]), r[1].strip()[:100], g("Instructions Executed"), g("# Samples"), g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Ideal")))
ti = sum(x[3] for x in rows); ts = sum(x[4] for x in rows); tw = sum(x[5] for x in rows)
print(f"total warp-inst {ti:.4g}  samples {ts:.0f}  smem wavefronts {tw:.4g}")
for f, ln, src, i, s, w, wi in sorted(rows, key=lambda x: -x[4])[:top]:
    print(f"{100*i/ti:5.1f}%i {100*s/max(ts,1):5.1f}%s wf {100*w/max(tw,1):5.1f}% (x{w/max(wi,1):.2f})  {f}:{ln:<4d} {src}")
