mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1c_smi.txt
nproc >> gpurun_out/r1c_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1c_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1c_pytest_gpu.log
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_$wl.json 2> gpurun_out/r1c_bench_$wl.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r1c_bench_$wl.json"))
print("$wl", round(d["value"],1), "GB/s e2e", round(d["e2e"]["value"],1), d["roofline"]["step_breakdown_ms"], d["gpu_launches"])
PY
done
bash tools/gpu_profile.sh r1c cfg2 14 > gpurun_out/r1c_profile.log 2>&1
tail -5 gpurun_out/r1c_profile.log
