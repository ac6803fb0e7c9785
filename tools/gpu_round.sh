#!/bin/bash
# One GPU-box session: parity tests, smoke, both bench arms, then the ncu launch list and one full capture of the
# dominant kernel (B200_PROFILING.md recipe).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
python bench.py --workload cfg3 --no-cpu > gpurun_out/bench_cfg3.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_cfg3.json
python bench.py --workload cfg3 --no-cpu --impl reference > gpurun_out/bench_cfg3_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_cfg3_ref.json
if [ "${1:-}" = "ncu" ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:partition_kernel -s 6 -c 2 -f -o gpurun_out/prof_partition \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:local_sort_kernel -s 3 -c 1 -f -o gpurun_out/prof_local \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/ncu_full.log 2>&1
fi
echo done
