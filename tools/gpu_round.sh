#!/bin/bash
# One GPU-box session: parity tests, smoke, both bench arms, the reference's own drivers on libb200sort.so, then the ncu
# launch list and one full capture (B200_PROFILING.md recipe).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
for w in cfg1 cfg3 cfg4; do
  python bench.py --workload $w --no-cpu > gpurun_out/bench_$w.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_$w.json
  python bench.py --workload $w --no-cpu --impl reference > gpurun_out/bench_${w}_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_${w}_ref.json
done
if [ -x oracle/_ref/lsb_sort_on_b200sort ]; then
  LD_LIBRARY_PATH=gpu_sort_b200 oracle/_ref/lsb_sort_on_b200sort --n=268435456 --t=3 > gpurun_out/ref_lsb_driver_on_b200sort.log 2>&1; tail -4 gpurun_out/ref_lsb_driver_on_b200sort.log
  LD_LIBRARY_PATH=gpu_sort_b200 oracle/_ref/msb_test_on_b200sort > gpurun_out/ref_msb_driver_on_b200sort.log 2>&1; tail -3 gpurun_out/ref_msb_driver_on_b200sort.log
fi
if [ -x oracle/_ref/msb_gtests_on_b200sort ]; then
  LD_LIBRARY_PATH=gpu_sort_b200 timeout 600 oracle/_ref/msb_gtests_on_b200sort --gtest_filter='Sort_Keys.Entropy_*:Sort_Pairs.*' -k 200000 -p 100000 > gpurun_out/ref_gtests_on_b200sort.log 2>&1; grep -E "PASSED|FAILED|tests ran" gpurun_out/ref_gtests_on_b200sort.log | tail -5
fi
if [ "${1:-}" = "ncu" ]; then
  # bounded captures: tools/one_sort.py runs 2 sorts; the launch list sees both, the full capture only the second sort's
  # data-moving kernels (skip the first sort's launches of the same families: 4 levels x (hist + scatter) + 4 on-chip = 12)
  for w in cfg2 cfg3; do
    python tools/one_sort.py $w > gpurun_out/plain_$w.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$w.csv python tools/one_sort.py $w > gpurun_out/ncu_launches_$w.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:"scatter|local_sort_kernel|tile_hist_kernel" -s 12 -c 12 -f -o gpurun_out/prof_$w \
        python tools/one_sort.py $w > gpurun_out/ncu_full_$w.log 2>&1
  done
fi
echo done
