#!/bin/bash
# Round-2 final evidence on one B200 (every step under its own timeout; most important first).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log | head -1
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 150 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench.json
timeout 100 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
: > gpurun_out/skew_ours_final.jsonl
for spec in "entropy 2" "entropy 3" "zipf_hash 0" "uniform 0"; do
  set -- $spec
  timeout 60 python bench.py --workload cfg4 --dist $1 --param $2 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras 2>/dev/null | cut -c1-400 >> gpurun_out/skew_ours_final.jsonl
done
cut -c1-160 gpurun_out/skew_ours_final.jsonl
timeout 60 python bench.py --workload cfg3 --no-cpu --no-e2e --impl reference > gpurun_out/bench_cfg3_ref.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_cfg3_ref.json
timeout 40 python tools/one_sort.py cfg2 > gpurun_out/plain_cfg2.log 2>&1 &&
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg2.csv python tools/one_sort.py cfg2 > gpurun_out/ncu_launches_cfg2.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"scatter|rank_sort|tile_hist_kernel" -s 4 -c 5 -f -o gpurun_out/prof_cfg2 python tools/one_sort.py cfg2 > gpurun_out/ncu_full_cfg2.log 2>&1
echo done
