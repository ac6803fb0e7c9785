#!/bin/bash
# Round-2 state check on one B200: per-kernel timings first (cheap), then bench both arms, the parity suite, ncu captures of cfg2.
set -u
mkdir -p gpurun_out
PERF_PROF=1 timeout 300 python tools/perf.py 28 7 msb32,lsb32v4,msb64,lsb64 > gpurun_out/perf_uniform.jsonl 2>&1; cat gpurun_out/perf_uniform.jsonl | cut -c1-900
PERF_PROF=1 timeout 200 python tools/perf.py 29 3 msb64 zipf_hash > gpurun_out/perf_zipf.jsonl 2>&1; cut -c1-900 gpurun_out/perf_zipf.jsonl
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json
timeout 200 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 60 python tools/one_sort.py cfg2 > gpurun_out/plain_cfg2.log 2>&1 &&
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg2.csv python tools/one_sort.py cfg2 > gpurun_out/ncu_launches_cfg2.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"scatter|local_sort_kernel|rank_sort|tile_hist_kernel" -s 8 -c 10 -f -o gpurun_out/prof_cfg2 python tools/one_sort.py cfg2 > gpurun_out/ncu_full_cfg2.log 2>&1
echo done
