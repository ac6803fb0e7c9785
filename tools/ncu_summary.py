#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box):  python tools/ncu_summary.py REP [more metrics...]  -> one block per kernel launch."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"] + sys.argv[2:]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("----")
    for w in want:
        if w in d:
            print(f"{w:75s} {d[w]:>18s} {units[hdr.index(w)]}")
    stalls = []
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") or ("warp_latency_issue_stalled" in k and k.endswith(".ratio")):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v > 0.2:
                stalls.append((v, k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "").replace(".ratio", "")))
    print("   stalls (warp-cycles per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)))
