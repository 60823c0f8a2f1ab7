#!/bin/bash
# tools/gpu_multi8.sh TAG N -- N GPUs of one box, launched as the driver does:
#   the default line (cfg2, weak scaling), BASELINE config 5 at its literal size (N x 12.7 GB of
#   FASTQ-like records = 100 GB at N = 8), the reference arm.
TAG=${1:-r1r}; N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt; nproc >> $OUT/${TAG}_gpus.txt; free -g | head -2 >> $OUT/${TAG}_gpus.txt
run() {  # name, port, args...
  name=$1; port=$2; shift; shift
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N "$@" ) > $OUT/${TAG}_${name}_n$N.json 2> $OUT/${TAG}_${name}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_${name}_n$N.json").read().strip().splitlines()[-1])
    print("$name", d["n_gpus"], "GPUs:", round(d["value"],1), "GB/s  ms/step", round(d["ms_per_step"],3), "e2e", d.get("e2e") and round(d["e2e"]["value"],1), "total_lines", d.get("config",{}).get("total_lines"), "total_records", d.get("config",{}).get("total_records"))
except Exception as e:
    print("$name failed", e)
PY
  grep real $OUT/${TAG}_${name}_n$N.err
}
run bench 29511 --steps 20 --warmup 6 --no-cpu-baseline
if [ "$N" = "8" ]; then
run cfg5_100GB 29512 --workload cfg5 --reads 39800000 --steps 3 --warmup 3 --no-cpu-baseline
run ref 29513 --impl reference --steps 2 --warmup 1
fi
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 6 --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench_n1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print("1 GPU on the same box:", round(d["value"],1), "GB/s")
PY
