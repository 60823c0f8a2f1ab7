#!/bin/bash
# tools/gpu_round10.sh TAG -- A/B of register-trimmed builds of the multi-part matcher (4 CTAs per SM)
TAG=${1:-r1y}
OUT=gpurun_out
mkdir -p $OUT
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2", round(d["value"],1), "GB/s  ms/step", round(d["ms_per_step"],4), {k: round(v,3) for k,v in d["roofline"]["kernels_ms"].items()})
except Exception as e:
    print("$2 failed", e)
PY
}
for wl in cfg3 cfg4; do
  for v in base va vb; do
    lib=""; [ $v != base ] && lib=seeq_b200/libseeq_b200_$v.so
    SEEQ_B200_LIB=$lib timeout 300 python bench.py --workload $wl --steps 10 --warmup 6 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${wl}_$v.json 2> $OUT/${TAG}_${wl}_$v.err
    show $OUT/${TAG}_${wl}_$v.json ${wl}_$v; tail -2 $OUT/${TAG}_${wl}_$v.err
  done
done
