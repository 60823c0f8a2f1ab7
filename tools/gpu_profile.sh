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
# gpurun copies back at most 64 MiB: summarise here (tools/summarise_ncu.py only needs `ncu -i`), keep the per-instruction
# source page of the largest kernels as CSV, and drop the report itself when it is too large to travel
python tools/summarise_ncu.py $OUT/${TAG}_${WL}_full.ncu-rep $OUT/${TAG}_${WL} --launches $OUT/${TAG}_${WL}_launches.csv \
    --note "ncu --set full --clock-control none of one step of bench.py --workload $WL (tools/gpu_profile.sh $TAG)" > /dev/null 2>&1
for k in k12_scan_pack k2_bitslice; do
  ncu -i $OUT/${TAG}_${WL}_full.ncu-rep --page source --csv --kernel-name regex:$k 2>/dev/null | python -c "
import csv, sys
keep = None
w = csv.writer(sys.stdout)
for r in csv.reader(sys.stdin):
    if r and r[0] == 'Address':
        want = ('Address', 'Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed', 'L1 Wavefronts Shared',
                'stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_math', 'stall_mio', 'stall_wait', 'stall_not_selected',
                'stall_selected', 'stall_branch_resolving', 'stall_no_inst', 'stall_lg')
        keep = [i for i, h in enumerate(r) if h in want]
    w.writerow([r[i] for i in keep if i < len(r)] if keep and len(r) > 8 else r)
" | gzip > $OUT/${TAG}_${WL}_${k}_source.csv.gz
done
SZ=$(du -sm $OUT | cut -f1)
if [ "$SZ" -gt 55 ]; then rm -f $OUT/${TAG}_${WL}_full.ncu-rep; fi
ls -la $OUT | tail -8
