#!/bin/bash
# tools/gpu_round.sh TAG -- one GPU session (run under gpurun): smoke, GPU parity
# suite, bench of the five BASELINE configs, launch list + ncu --set full of cfg2.
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt
nproc >> gpurun_out/${TAG}_smi.txt
if ! timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; then
  echo "SMOKE FAILED"; tail -20 gpurun_out/${TAG}_smoke.log; exit 1
fi
tail -1 gpurun_out/${TAG}_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -15 gpurun_out/${TAG}_pytest_gpu.log
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$wl.json"))
    print("$wl", round(d["value"],1), "GB/s e2e", round(d["e2e"]["value"],1), d["roofline"]["step_breakdown_ms"], d["gpu_launches"])
except Exception as e:
    print("$wl failed", e)
PY
done
# the two lines the driver takes at round end: our arm with defaults, the reference arm
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 600 gpurun_out/${TAG}_bench_reference.json; grep real gpurun_out/${TAG}_bench_default.err gpurun_out/${TAG}_bench_reference.err
bash tools/gpu_profile.sh $TAG cfg2 11 > gpurun_out/${TAG}_profile.log 2>&1
tail -3 gpurun_out/${TAG}_profile.log
bash tools/gpu_profile.sh $TAG cfg3 0 >> gpurun_out/${TAG}_profile.log 2>&1
bash tools/gpu_profile.sh $TAG cfg4 0 >> gpurun_out/${TAG}_profile.log 2>&1
