#!/bin/bash
# tools/gpu_round6.sh TAG -- pattern sets: parity tests + throughput against separate scans
TAG=${1:-r1q}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=4 > $OUT/${TAG}_pytest_multi.log 2>&1; tail -40 $OUT/${TAG}_pytest_multi.log
timeout 600 python tools/bench_multi.py > $OUT/${TAG}_bench_multi.jsonl 2> $OUT/${TAG}_bench_multi.err; cat $OUT/${TAG}_bench_multi.jsonl | cut -c1-420; tail -3 $OUT/${TAG}_bench_multi.err
