#!/bin/bash
# tools/gpu_round9.sh TAG -- SQB_TIMING inside the replayed graph: cost against timing outside the timed region
TAG=${1:-r1v}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_large.py -m gpu -q -k "leader or graph" > $OUT/${TAG}_pytest.log 2>&1; tail -5 $OUT/${TAG}_pytest.log
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2", round(d["value"],1), "GB/s  ms/step", round(d["ms_per_step"],4), {k: round(v,3) for k,v in d["roofline"]["kernels_ms"].items()}, d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["gpu_launches"])
except Exception as e:
    print("$2 failed", e)
PY
}
for wl in cfg2 cfg1 cfg5; do
  for t in 1 0; do
    SEEQ_B200_BENCH_TIMING=$t timeout 300 python bench.py --workload $wl --steps 20 --warmup 6 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${wl}_t$t.json 2> $OUT/${TAG}_${wl}_t$t.err
    show $OUT/${TAG}_${wl}_t$t.json ${wl}_timing$t; tail -2 $OUT/${TAG}_${wl}_t$t.err
  done
done
SEEQ_B200_GRAPHS=0 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 6 --no-cpu-baseline --no-e2e > $OUT/${TAG}_cfg2_nograph.json 2> $OUT/${TAG}_cfg2_nograph.err
show $OUT/${TAG}_cfg2_nograph.json cfg2_timing1_nograph
