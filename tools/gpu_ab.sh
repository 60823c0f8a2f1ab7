#!/bin/bash
# tools/gpu_ab.sh TAG -- GPU parity suite + A/B runs of tuning knobs
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -5 $OUT/${TAG}_pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload ${WL:-cfg2} --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_${name}.json"))
    print("${WL:-cfg2} $name", round(d["value"],1), "GB/s", {k: round(v,4) for k,v in d["roofline"]["kernels_ms"].items()}, {k: round(v,4) for k,v in d["roofline"]["step_breakdown_ms"].items()})
except Exception as e:
    print("$name failed", e)
PY
}
WL=cfg2 run cfg2_nfa X=1
WL=cfg2 run cfg2_myers SEEQ_B200_NFA=0
WL=cfg1 run cfg1_nfa X=1
WL=cfg1 run cfg1_myers SEEQ_B200_NFA=0
WL=cfg5 run cfg5_nfa X=1
WL=cfg5 run cfg5_myers SEEQ_B200_NFA=0
WL=cfg2 run cfg2_filter SEEQ_B200_FILTER=2
ls $OUT | wc -l
