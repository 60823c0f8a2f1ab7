#!/bin/bash
# tools/gpu_multi.sh TAG N -- bench.py on N GPUs of one box exactly as the driver launches it
TAG=${1:-mg}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -c 1500 gpurun_out/${TAG}_bench_n$N.json; tail -5 gpurun_out/${TAG}_bench_n$N.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 ) > gpurun_out/${TAG}_ref_n$N.json 2> gpurun_out/${TAG}_ref_n$N.err
tail -c 600 gpurun_out/${TAG}_ref_n$N.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2>/dev/null
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_n%d.json"%n).read().strip().splitlines()[-1])
        print(n, "GPUs:", round(d["value"],1), "GB/s  e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],3), "total_lines", d["config"]["total_lines"], "total_records", d["config"]["total_records"])
    except Exception as e:
        print(n, "failed", e)
PY
