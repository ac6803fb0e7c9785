#!/bin/bash
# Multi-GPU evidence on one box with N GPUs: driver-style bench (both arms) + skew checks of the exchange path.  usage: tools/gpu_dist_n.sh N [quick]
N=${1:-2}; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; cut -c1-1800 gpurun_out/bench_n$N.json
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --impl reference > gpurun_out/bench_n${N}_ref.json 2>> gpurun_out/bench_n$N.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_n${N}_ref.json
: > gpurun_out/dist_check_n$N.jsonl
dists="uniform constant zipf_hash entropy sorted"; [ "${2:-}" = "quick" ] && dists="uniform zipf_hash constant"
for d in $dists; do
  timeout 200 $TR tools/dist_check.py 26 1 $d 3 2>> gpurun_out/bench_n$N.err | cut -c1-400 | tee -a gpurun_out/dist_check_n$N.jsonl
done
timeout 200 $TR tools/dist_check.py 26 0 uniform 3 2>> gpurun_out/bench_n$N.err | cut -c1-400 | tee -a gpurun_out/dist_check_n$N.jsonl
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/bench_n$N.err | tail -5
