#!/bin/bash
# tools/gpu_round8.sh TAG -- A/B of k15_pack at 4 / 5 / 6 CTAs per SM (three builds of the library)
TAG=${1:-r1t}
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
for wl in cfg2 cfg5; do
  for v in 4 3 2; do
    lib=""; [ $v != 4 ] && lib=seeq_b200/libseeq_b200_p$v.so
    SEEQ_B200_LIB=$lib timeout 300 python bench.py --workload $wl --steps 10 --warmup 6 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${wl}_p$v.json 2> $OUT/${TAG}_${wl}_p$v.err
    show $OUT/${TAG}_${wl}_p$v.json ${wl}_pack$v; tail -2 $OUT/${TAG}_${wl}_p$v.err
  done
done
