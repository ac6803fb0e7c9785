mkdir -p gpurun_out
PERF_PROF=1 python tools/perf.py 28 5 msb32,lsb32v4,msb64 2>&1 | cut -c1-760
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
