#!/bin/bash
# tools/gpu_round3.sh TAG -- GPU parity suite, graph replay on/off on three workloads, ncu of K1 on the FASTQ-like input
TAG=${1:-r1n}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -14 $OUT/${TAG}_pytest_gpu.log
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
for wl in cfg1 cfg2 cfg5; do
  for g in 1 0; do
    SEEQ_B200_GRAPHS=$g timeout 300 python bench.py --workload $wl --steps 20 --warmup 4 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${wl}_graphs$g.json 2> $OUT/${TAG}_${wl}_graphs$g.err
    show $OUT/${TAG}_${wl}_graphs$g.json ${wl}_graphs$g; tail -2 $OUT/${TAG}_${wl}_graphs$g.err
  done
done
for wl in cfg3 cfg4; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 4 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${wl}.json 2> $OUT/${TAG}_${wl}.err
  show $OUT/${TAG}_${wl}.json ${wl}; tail -2 $OUT/${TAG}_${wl}.err
done
timeout 600 python bench.py --workload cfg5 --reads 39800000 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cfg5_12GB.json 2> $OUT/${TAG}_bench_cfg5_12GB.err
show $OUT/${TAG}_bench_cfg5_12GB.json cfg5_12GB; tail -3 $OUT/${TAG}_bench_cfg5_12GB.err
BENCH="python bench.py --workload cfg5 --steps 2 --warmup 4 --no-cpu-baseline --no-e2e"
SEEQ_B200_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_scan_classify' -s 4 -c 1 -f -o $OUT/${TAG}_cfg5_k1 $BENCH > $OUT/${TAG}_cfg5_k1.log 2>&1
ls -la $OUT/${TAG}_cfg5_k1* | tail -3
