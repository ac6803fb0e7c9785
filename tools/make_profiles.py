#!/usr/bin/env python
"""Turns the ncu outputs a GPU session left in gpurun_out/ into the tracked summaries under profiles/ (run on the CPU box).

  launches.csv (ncu --metrics gpu__time_duration.sum)   -> profiles/rNN_launches_<tag>.txt   per-kernel share of the step
  prof_<tag>.ncu-rep (ncu --set full)                   -> profiles/rNN_full_<tag>.txt        + profiles/traffic.json
usage: make_profiles.py ROUND TAG WORKLOAD SORTS_CAPTURED   e.g.  make_profiles.py r01 cfg2 cfg2 5
"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd, tag, workload, sorts = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
import re
def family(name):
    """kernel family as bench.py / b200_prof_report name it, from a (possibly abbreviated) demangled kernel name"""
    args = [x.strip().split(")")[-1] for x in re.sub(r"^[^<]*<", "", name).split(">")[0].split(",")]
    if "scatter_stable_fast_kernel" in name: return "scatter_stable"
    if "scatter_fast_kernel" in name: return "scatter"
    if "scatter_kernel" in name and len(args) >= 7:
        mode, ord_ = args[5], args[6]
        return {"0": "scatter_stable" if ord_ in ("1", "true") else "scatter", "1": "scatter_onesweep", "2": "range_partition"}.get(mode, "scatter")
    if "rank_sort_kernel" in name: return "local_sort_rank_dense" if (len(args) >= 7 and args[6] in ("1", "true")) else "local_sort_rank"
    if "bitmap_sort_kernel" in name: return "local_sort_bitmap"
    if "local_sort_kernel" in name and len(args) >= 6:
        return "local_sort_count" if args[4] == "1" else "local_sort_lsd"
    if "tile_hist_kernel" in name: return "tile_hist"
    if "hist_all_kernel" in name: return "hist_all"
    return None

lc = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = None; agg = collections.OrderedDict()
    for r in rows:
        if r[0] == "ID": hdr = r; continue
        if hdr is None: continue
        d = dict(zip(hdr, r)); name = d["Kernel Name"]
        try: v = float(d["Metric Value"].replace(",", ""))
        except ValueError: continue
        u = d["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v * 1e6 if u.startswith("s") else v      # -> us
        a = agg.setdefault(name[:110], [0, 0.0]); a[0] += 1; a[1] += v
    ours = {k: v for k, v in agg.items() if "b200::" in k}
    tot = sum(v[1] for v in ours.values())
    with open(os.path.join(ROOT, "profiles", f"{rnd}_launches_{tag}.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py workload {workload}; per-launch times are cold-cache and\n"
                f"# serialised: compare SHARES with bench.py's CUDA-event figures (roofline.kernels_ms_per_step), not absolutes.\n")
        for k, (c, t) in agg.items():
            share = f"{100 * t / tot:5.1f}% of the library's kernels" if k in ours else "(harness)"
            f.write(f"{c:5d} launches {t:12.1f} us  {share}  {k}\n")
    print(open(os.path.join(ROOT, "profiles", f"{rnd}_launches_{tag}.txt")).read())

rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr = rows[0]; units = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    def val(r, k):
        v = float(r[ix[k]].replace(",", "")); u = units[ix[k]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1}.get(u, 1)
    fam = collections.defaultdict(list)
    for r in rows[2:]:
        f = family(r[ix["Kernel Name"]])
        if f: fam[f].append((val(r, "gpu__time_duration.sum"), val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")))
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tj)) if os.path.exists(tj) else {}
    with open(os.path.join(ROOT, "profiles", f"{rnd}_full_{tag}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, workload {workload}, {sorts} sorts captured; per kernel family:\n")
        for k, L in fam.items():
            big = [x for x in L if x[0] > 50e-6]
            per_step = sum(b for _, b in big) / sorts
            traffic.setdefault(k, {})[workload] = int(per_step)
            f.write(f"{k:18s} launches {len(L):4d} (of which > 50 us: {len(big)})  dram bytes per sort {per_step/1e9:8.3f} GB  time per sort {sum(t for t,_ in big)/sorts*1e3:7.3f} ms (under ncu)\n")
        f.write("\n")
        f.write(subprocess.run(["python", os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
    json.dump(traffic, open(tj, "w"), indent=1)
    print(json.dumps(traffic))
