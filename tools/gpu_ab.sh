#!/bin/bash
# tools/gpu_ab.sh TAG -- A/B runs of tuning knobs on cfg2 + ncu of the matcher on cfg3 / cfg4
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu.log
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
run base X=1
run l2f32 SEEQ_B200_L2_FETCH=32
run l2f128 SEEQ_B200_L2_FETCH=128
run chunk1 SEEQ_B200_REV_CHUNK=1
run chunk4 SEEQ_B200_REV_CHUNK=4
run chunk1_l2f32 SEEQ_B200_REV_CHUNK=1 SEEQ_B200_L2_FETCH=32
run nocuts SEEQ_B200_CUTS=0
WL=cfg4 run cfg4_base X=1
WL=cfg4 run cfg4_ctas3 SEEQ_B200_BS_CTAS=3
WL=cfg3 run cfg3_base X=1
for wl in cfg4 cfg3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k2_bitslice|k15_pack|k34_' -s 6 -c 3 \
      -f -o $OUT/${TAG}_${wl}_full python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${wl}_full.log 2>&1
done
ls -la $OUT | tail -5
