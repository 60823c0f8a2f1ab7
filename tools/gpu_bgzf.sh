#!/bin/bash
# tools/gpu_bgzf.sh TAG -- run on the GPU box (under gpurun): the BGZF tests, the BGZF bench line, one ncu capture of
# the inflate kernel.  Outputs land in gpurun_out/.
TAG=${1:-r6a}
OUT=gpurun_out
mkdir -p $OUT
timeout 150 python -m pytest tests/test_gpu_bgzf.py -q -p no:cacheprovider > $OUT/${TAG}_pytest_bgzf.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_bgzf.log
timeout 120 python tools/bgzf_bench.py --out $OUT/${TAG}_bgzf.json > $OUT/${TAG}_bgzf.log 2>&1
echo "bench rc=$?"; tail -c 1500 $OUT/${TAG}_bgzf.log
if [ "$2" != "noncu" ]; then
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k0_inflate -c 1 -f -o $OUT/${TAG}_bgzf_full \
    python tools/bgzf_bench.py --inflate-only > $OUT/${TAG}_bgzf_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i $OUT/${TAG}_bgzf_full.ncu-rep --page raw --csv > $OUT/${TAG}_bgzf_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_bgzf_full.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/${TAG}_bgzf_source.csv.gz
fi
if [ "$3" = "bench" ]; then
timeout 150 python bench.py --no-cpu-baseline --no-configs --steps 10 --warmup 6 > $OUT/${TAG}_bench_quick.json 2> $OUT/${TAG}_bench_quick.err
echo "bench.py rc=$?"; python -c "
import json,sys
d=json.loads(open('$OUT/${TAG}_bench_quick.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'bgzf',json.dumps(d['bgzf'])[:900])"
fi
ls -la $OUT | grep ${TAG}
if [ "$4" = "sanitize" ]; then
for tool in memcheck racecheck; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout 55 compute-sanitizer --tool $tool $extra --print-limit 20 python tools/sanitize_bgzf.py > $OUT/${TAG}_sanitize_bgzf_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize bgzf tour|MISMATCH" $OUT/${TAG}_sanitize_bgzf_$tool.log | tail -4
done
fi
