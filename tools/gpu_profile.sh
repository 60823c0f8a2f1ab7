#!/bin/bash
# tools/gpu_profile.sh TAG [WORKLOAD] -- run on the GPU box (under gpurun):
#   1. launch list of one short bench run (per-launch device time, cold cache, serialised)
#   2. ncu --set full of one launch of every kernel of the step
# Outputs land in gpurun_out/ and are summarised into profiles/ by tools/summarise_ncu.py.
TAG=${1:-r1}
WL=${2:-cfg2}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_${WL}_launches.csv $BENCH > $OUT/${TAG}_${WL}_launches.log 2>&1
# skip the generator + 3 warm-up steps, then capture one full step
NK=${3:-14}
if [ "$NK" = "0" ]; then exit 0; fi        # launch list only
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_|k12_|k15_|k2_|k34_|k_tile|k_off' \
    -s $((3 * NK)) -c $NK -f -o $OUT/${TAG}_${WL}_full $BENCH > $OUT/${TAG}_${WL}_full.log 2>&1
ls -la $OUT | tail -8
