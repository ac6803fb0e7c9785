#!/usr/bin/env python
"""Dynamic opcode mix + hottest instructions of the first kernel in an .ncu-rep (needs --import-source on).
usage: ncu_ops.py REP [top=25]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; ops = collections.Counter(); smp = collections.Counter(); lines = []
kern = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kern += 1
        if kern > 1: break
        print("#", r[1]); continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    d = dict(zip(hdr, r))
    try: n = float(d["Instructions Executed"]); s = float(d["# Samples"])
    except (ValueError, KeyError): continue
    src = d["Source"].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.rstrip(";")
    ops[op.split(".")[0]] += n; smp[op.split(".")[0]] += s
    lines.append((s, n, src, d))
tot = sum(ops.values()); tots = sum(smp.values())
print(f"# {tot:.4g} warp-instructions, {tots:.0f} samples")
for op, n in ops.most_common(22): print(f"{op:10s} {100*n/tot:5.1f}% of instr  {100*smp[op]/max(tots,1):5.1f}% of samples")
print("# hottest by stall samples")
stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for s, n, src, d in sorted(lines, key=lambda x: -x[0])[:top]:
    st = sorted(((float(d[k] or 0), k[6:]) for k in stall_keys), reverse=True)[:3]
    print(f"{100*s/tots:5.2f}%s {100*n/tot:5.2f}%i {src[:60]:60s} " + " ".join(f"{k}={v:.0f}" for v, k in st if v > 0))
