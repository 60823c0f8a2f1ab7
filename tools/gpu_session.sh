#!/bin/bash
# tools/gpu_session.sh TAG PART... -- one GPU session under gpurun; every part writes into gpurun_out/TAG_*.
# (The numbered sessions of round 1, r1m .. r2b, were runs of these parts in various combinations.)
#
#   tests              smoke() + the GPU parity suite
#   bench              bench.py on the five BASELINE configs (device-resident + end to end)
#   big                config 5 at one GPU's share of 100 GB (39.8 M records = 12.7 GB, sqbScanDeviceLarge)
#   lines              the two lines the driver takes: default arm and `--impl reference`
#   profile [WL]       launch list + `ncu --set full` of one step (default cfg2) -> tools/summarise_ncu.py
#   multi              pattern sets: throughput against separate scans (tools/bench_multi.py)
#   h2d                bare pinned H2D copy against sqbScanHost at several chunk sizes (tools/bench_h2d.py)
#   ab WL LIB...       bench.py on workload WL with the stock library and every LIB (another build of the
#                      library, e.g. nvcc ... -DSQB_PACK_CTAS=5 -> seeq_b200/libseeq_b200_p5.so; SEEQ_B200_LIB)
#   sanitize           compute-sanitizer memcheck + racecheck over the reduced kernel tour (tools/sanitize_case.py)
#   bgzf               the BGZF (bgzip) input path: tools/gpu_bgzf.sh (tests, bench line, ncu, sanitizers)
#   scale N            (gpurun --gpus N) torchrun bench.py on N GPUs as the driver launches it, plus config 5 at
#                      N x 12.7 GB and the reference arm when N = 8, plus one GPU of the same box
TAG=${1:?tag}; shift
OUT=gpurun_out
mkdir -p $OUT
show() {  # file label
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$2", d.get("n_gpus"), "GPU:", round(d["value"],1), "GB/s  ms/step", round(d["ms_per_step"],4), "e2e", d.get("e2e") and round(d["e2e"]["value"],1),
          {k: round(v["ms"],3) for k,v in (r.get("kernels") or {}).items()}, {k: round(v,3) for k,v in (r.get("step_breakdown_ms") or {}).items()},
          "launches", d.get("gpu_launches"), "reruns", d.get("scan_reruns"))
except Exception as e:
    print("$2 failed", e)
PY
}
while [ $# -gt 0 ]; do
  part=$1; shift
  case $part in
  tests)
    nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_smi.txt; nproc >> $OUT/${TAG}_smi.txt; free -g >> $OUT/${TAG}_smi.txt
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $OUT/${TAG}_smoke.log; }
    tail -1 $OUT/${TAG}_smoke.log
    timeout 1700 python -m pytest tests -m gpu -q --durations=4 > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -12 $OUT/${TAG}_pytest_gpu.log ;;
  bench)
    for wl in metric cfg1 cfg2 cfg3 cfg4 cfg5; do
      timeout 300 python bench.py --workload $wl --steps 20 --warmup 6 --no-cpu-baseline --no-configs > $OUT/${TAG}_bench_$wl.json 2> $OUT/${TAG}_bench_$wl.err
      show $OUT/${TAG}_bench_$wl.json $wl; tail -2 $OUT/${TAG}_bench_$wl.err
    done ;;
  big)
    timeout 600 python bench.py --workload cfg5 --reads 39800000 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $OUT/${TAG}_bench_cfg5_12GB.json 2> $OUT/${TAG}_bench_cfg5_12GB.err
    show $OUT/${TAG}_bench_cfg5_12GB.json cfg5_12GB; tail -3 $OUT/${TAG}_bench_cfg5_12GB.err ;;
  lines)
    ( time timeout 900 python bench.py ) > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
    show $OUT/${TAG}_bench_default.json default; grep real $OUT/${TAG}_bench_default.err
    ( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
    tail -c 500 $OUT/${TAG}_bench_reference.json; grep real $OUT/${TAG}_bench_reference.err ;;
  profile)
    wl=cfg2; case "$1" in cfg*|metric) wl=$1; shift ;; esac
    SEEQ_B200_GRAPHS=0 bash tools/gpu_profile.sh $TAG $wl 11 > $OUT/${TAG}_profile.log 2>&1; tail -3 $OUT/${TAG}_profile.log ;;
  multi)
    timeout 600 python tools/bench_multi.py > $OUT/${TAG}_bench_multi.jsonl 2> $OUT/${TAG}_bench_multi.err; cut -c1-420 $OUT/${TAG}_bench_multi.jsonl; tail -3 $OUT/${TAG}_bench_multi.err ;;
  h2d)
    timeout 600 python tools/bench_h2d.py > $OUT/${TAG}_h2d.json 2> $OUT/${TAG}_h2d.err; cat $OUT/${TAG}_h2d.json; tail -3 $OUT/${TAG}_h2d.err ;;
  ab)
    wl=$1; shift
    for lib in "" "$@"; do
      name=stock; [ -n "$lib" ] && name=$(basename $lib .so)
      SEEQ_B200_LIB=$lib timeout 300 python bench.py --workload $wl --steps 10 --warmup 6 --no-cpu-baseline --no-e2e --no-configs > $OUT/${TAG}_${wl}_$name.json 2> $OUT/${TAG}_${wl}_$name.err
      show $OUT/${TAG}_${wl}_$name.json ${wl}_$name; tail -2 $OUT/${TAG}_${wl}_$name.err
    done
    set -- ;;
  scale)
    N=$1; shift
    nvidia-smi -L > $OUT/${TAG}_gpus.txt; nproc >> $OUT/${TAG}_gpus.txt; free -g | head -2 >> $OUT/${TAG}_gpus.txt
    run() {  # name port args...
      name=$1; port=$2; shift; shift
      ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
          bench.py --gpus $N "$@" ) > $OUT/${TAG}_${name}_n$N.json 2> $OUT/${TAG}_${name}_n$N.err
      show $OUT/${TAG}_${name}_n$N.json $name; grep real $OUT/${TAG}_${name}_n$N.err
    }
    run bench 29511 --steps 20 --warmup 6 --no-cpu-baseline
    if [ "$N" = "8" ]; then
      run cfg5_100GB 29512 --workload cfg5 --reads 39800000 --steps 3 --warmup 3 --no-cpu-baseline
      run ref 29513 --impl reference --steps 2 --warmup 1
    fi
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 6 --no-cpu-baseline --no-e2e --no-configs > $OUT/${TAG}_bench_n1.json 2>/dev/null
    show $OUT/${TAG}_bench_n1.json same_box ;;
  sanitize)
    bash tools/gpu_sanitize.sh $TAG memcheck racecheck ;;
  bgzf)
    # BGZF input: GPU tests, tools/bgzf_bench.py line, ncu of the inflate kernel, bench.py with the bgzf key, sanitizers
    bash tools/gpu_bgzf.sh $TAG ncu bench sanitize ;;
  *) echo "unknown part $part"; exit 2 ;;
  esac
done
