#!/bin/bash
# Config 4 (2^29 uint64 keys, MSB hybrid sort) over the key distributions SURVEY.md section 8d lists, both arms side by side.
# usage: tools/skew_sweep.sh [out=gpurun_out/skew_sweep.jsonl]
out=${1:-gpurun_out/skew_sweep.jsonl}; mkdir -p "$(dirname $out)"; : > $out
for spec in "uniform 0" "entropy 2" "entropy 3" "entropy 5" "zipf_hash 0" "zipf_rank 0" "sorted 0" "reverse 0" "constant 0"; do
  set -- $spec
  for impl in ours reference; do
    timeout 150 python bench.py --workload cfg4 --dist $1 --param $2 --impl $impl --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras 2>/dev/null \
      | python -c "import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'dist':'$1','param':$2,'impl':d.get('impl','ours'),'ms_median':d.get('ms_median'),'gkeys_s':d.get('value_median'),'verified':d.get('verified')}))" >> $out || echo "{\"dist\":\"$1\",\"param\":$2,\"impl\":\"$impl\",\"failed\":true}" >> $out
  done
done
cat $out
