#!/bin/bash
# tools/gpu_round4.sh TAG -- full GPU suite (no -x), the five configs + the 12.7 GB shard, launch lists
TAG=${1:-r1o}
OUT=gpurun_out
mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -q --durations=6 > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -30 $OUT/${TAG}_pytest_gpu.log
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2", round(d["value"],1), "GB/s  ms/step", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d["roofline"]["kernels_ms"].items()}, {k: round(v,3) for k,v in d["roofline"]["step_breakdown_ms"].items()}, d["gpu_launches"])
except Exception as e:
    print("$2 failed", e)
PY
}
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 4 --no-cpu-baseline > $OUT/${TAG}_bench_${wl}.json 2> $OUT/${TAG}_bench_${wl}.err
  show $OUT/${TAG}_bench_${wl}.json ${wl}; tail -2 $OUT/${TAG}_bench_${wl}.err
done
timeout 600 python bench.py --workload cfg5 --reads 39800000 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cfg5_12GB.json 2> $OUT/${TAG}_bench_cfg5_12GB.err
show $OUT/${TAG}_bench_cfg5_12GB.json cfg5_12GB; tail -3 $OUT/${TAG}_bench_cfg5_12GB.err
for wl in cfg2 cfg5; do
SEEQ_B200_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_${wl}_launches.csv python bench.py --workload $wl --steps 2 --warmup 4 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${wl}_launches.log 2>&1
tail -16 $OUT/${TAG}_${wl}_launches.csv | cut -d, -f2,5 | cut -c1-120
done
