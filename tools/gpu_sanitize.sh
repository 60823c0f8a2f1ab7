#!/bin/bash
# tools/gpu_sanitize.sh TAG [TOOLS...] -- compute-sanitizer over the reduced kernel tour (tools/sanitize_case.py)
# on the GPU box; logs -> gpurun_out/TAG_sanitize_<tool>.log (summaries are copied to profiles/).
TAG=${1:?tag}; shift
TOOLS=${@:-memcheck racecheck}
OUT=gpurun_out
mkdir -p $OUT
for tool in $TOOLS; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  SANITIZE_READS=${SANITIZE_READS:-12000} timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 20 \
      python tools/sanitize_case.py > $OUT/${TAG}_sanitize_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize tour|MISMATCH" $OUT/${TAG}_sanitize_$tool.log | tail -5
done
