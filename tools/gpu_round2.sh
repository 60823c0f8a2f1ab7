#!/bin/bash
# tools/gpu_round2.sh TAG -- GPU session: smoke, GPU parity suite (incl. the large / full-size tests),
# bench of the five configs, cfg5 at one GPU's share of the 100 GB read set (12.7 GB), default line.
TAG=${1:-r1m}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_smi.txt
nproc >> $OUT/${TAG}_smi.txt; free -g >> $OUT/${TAG}_smi.txt
for d in /sys/bus/pci/devices/*; do
  if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa $(cat $d/numa_node 2>/dev/null)" >> $OUT/${TAG}_smi.txt; fi
done
ls /sys/devices/system/node/ >> $OUT/${TAG}_smi.txt 2>&1
if ! timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; then
  echo "SMOKE FAILED"; tail -20 $OUT/${TAG}_smoke.log; exit 1
fi
tail -1 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -25 $OUT/${TAG}_pytest_gpu.log
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2", round(d["value"],1), "GB/s e2e", d["e2e"] and round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d["roofline"]["kernels_ms"].items()}, {k: round(v,3) for k,v in d["roofline"]["step_breakdown_ms"].items()}, d["gpu_launches"])
except Exception as e:
    print("$2 failed", e)
PY
}
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$wl.json 2> $OUT/${TAG}_bench_$wl.err
  show $OUT/${TAG}_bench_$wl.json $wl
done
timeout 600 python bench.py --workload cfg5 --reads 39800000 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cfg5_12GB.json 2> $OUT/${TAG}_bench_cfg5_12GB.err
show $OUT/${TAG}_bench_cfg5_12GB.json cfg5_12GB; tail -3 $OUT/${TAG}_bench_cfg5_12GB.err
( time timeout 900 python bench.py ) > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
show $OUT/${TAG}_bench_default.json default; grep real $OUT/${TAG}_bench_default.err
