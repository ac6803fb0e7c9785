for rm in 1537 1024 512 256; do
  echo "# RANK_MIN=$rm"
  for l in 24 25 26 27; do B200SORT_RANK_MIN=$rm python tools/perf.py $l 5 msb32,lsb32v4 2>&1 | cut -c1-110; done
done
