#!/bin/bash
# Round-end evidence on one B200, most important first (the GPU budget may cut the tail): parity suite, smoke, both bench arms,
# ncu launch list + one full capture for cfg2, then the other workloads and the cfg3 captures.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 120 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 120 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench.json
timeout 60 python bench.py --workload cfg3 --no-cpu > gpurun_out/bench_cfg3.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_cfg3.json
timeout 60 python bench.py --workload cfg3 --no-cpu --impl reference > gpurun_out/bench_cfg3_ref.json 2>> gpurun_out/bench.err
ncu_caps() {
  w=$1
  timeout 60 python tools/one_sort.py $w > gpurun_out/plain_$w.log 2>&1 &&
  timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$w.csv python tools/one_sort.py $w > gpurun_out/ncu_launches_$w.log 2>&1
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:"scatter|local_sort_kernel|tile_hist_kernel" -s 12 -c 12 -f -o gpurun_out/prof_$w \
      python tools/one_sort.py $w > gpurun_out/ncu_full_$w.log 2>&1
}
ncu_caps cfg2
for w in cfg4 cfg1; do
  timeout 60 python bench.py --workload $w --no-cpu > gpurun_out/bench_$w.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_$w.json
  timeout 90 python bench.py --workload $w --no-cpu --impl reference > gpurun_out/bench_${w}_ref.json 2>> gpurun_out/bench.err
done
ncu_caps cfg3
echo done
