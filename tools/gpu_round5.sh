#!/bin/bash
# tools/gpu_round5.sh TAG -- full GPU suite, all configs, the two lines the driver takes, ncu of the cfg2 step
TAG=${1:-r1x}
OUT=gpurun_out
mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -q --durations=4 > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -12 $OUT/${TAG}_pytest_gpu.log
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2", round(d["value"],1), "GB/s  ms/step", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d["roofline"]["kernels_ms"].items()}, {k: round(v,3) for k,v in d["roofline"]["step_breakdown_ms"].items()}, d["gpu_launches"], "reruns", d.get("scan_reruns"))
except Exception as e:
    print("$2 failed", e)
PY
}
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 6 --no-cpu-baseline > $OUT/${TAG}_bench_${wl}.json 2> $OUT/${TAG}_bench_${wl}.err
  show $OUT/${TAG}_bench_${wl}.json ${wl}; tail -2 $OUT/${TAG}_bench_${wl}.err
done
timeout 600 python bench.py --workload cfg5 --reads 39800000 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cfg5_12GB.json 2> $OUT/${TAG}_bench_cfg5_12GB.err
show $OUT/${TAG}_bench_cfg5_12GB.json cfg5_12GB; tail -3 $OUT/${TAG}_bench_cfg5_12GB.err
( time timeout 900 python bench.py ) > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
show $OUT/${TAG}_bench_default.json default; grep real $OUT/${TAG}_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
tail -c 500 $OUT/${TAG}_bench_reference.json; grep real $OUT/${TAG}_bench_reference.err
export SEEQ_B200_GRAPHS=0
bash tools/gpu_profile.sh $TAG cfg2 11 > $OUT/${TAG}_profile.log 2>&1
tail -3 $OUT/${TAG}_profile.log
