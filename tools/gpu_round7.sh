#!/bin/bash
# tools/gpu_round7.sh TAG -- A/B of the multi-part matcher at 3 / 4 CTAs per SM (two builds of the library),
# the (26,4) shape on cfg4, device-resident results of the large scan, the new GPU tests
TAG=${1:-r1s}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_large.py tests/test_gpu_multi.py -m gpu -q -k "not full_size" > $OUT/${TAG}_pytest_large.log 2>&1; tail -6 $OUT/${TAG}_pytest_large.log
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2", round(d["value"],1), "GB/s  ms/step", round(d["ms_per_step"],4), {k: round(v,3) for k,v in d["roofline"]["kernels_ms"].items()}, "reruns", d.get("scan_reruns"))
except Exception as e:
    print("$2 failed", e)
PY
}
for wl in cfg3 cfg4; do
  for lib in "" seeq_b200/libseeq_b200_g4.so; do
    tag=${wl}_g3; [ -n "$lib" ] && tag=${wl}_g4
    SEEQ_B200_LIB=$lib timeout 300 python bench.py --workload $wl --steps 10 --warmup 6 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${tag}.json 2> $OUT/${TAG}_${tag}.err
    show $OUT/${TAG}_${tag}.json $tag; tail -2 $OUT/${TAG}_${tag}.err
  done
done
timeout 600 python bench.py --workload cfg5 --reads 39800000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench_cfg5_12GB.json 2> $OUT/${TAG}_bench_cfg5_12GB.err
show $OUT/${TAG}_bench_cfg5_12GB.json cfg5_12GB; tail -3 $OUT/${TAG}_bench_cfg5_12GB.err
