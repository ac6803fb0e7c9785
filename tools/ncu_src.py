#!/usr/bin/env python
"""Top lines of an `ncu --page source --csv` dump by stall samples / executed instructions (first kernel in file)."""
import csv, sys
f = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(f)))
# find header rows ("Address","Source",...) ; take first block
blocks = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
start = blocks[which]; end = blocks[which + 1] - 1 if which + 1 < len(blocks) else len(rows)
print(rows[start - 1][:2])
hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
def num(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(num(r, "# Samples") for r in body); tot_i = sum(num(r, "Instructions Executed") for r in body)
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
print("== by samples")
for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:top]:
    st = {k: num(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    top2 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*num(r,'# Samples')/tot_s:5.1f}%  inst {100*num(r,'Instructions Executed')/tot_i:5.2f}%  {r[ix['Source']][:90]:90s} {top2}")
print("== by instructions")
for r in sorted(body, key=lambda r: -num(r, "Instructions Executed"))[:top]:
    print(f"inst {100*num(r,'Instructions Executed')/tot_i:5.2f}%  samples {100*num(r,'# Samples')/tot_s:5.1f}%  {r[ix['Source']][:100]}")
