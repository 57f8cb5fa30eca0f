#!/bin/bash
# compute-sanitizer passes over smoke() (one tiny forward through every kernel of the hot path)
# and over a single causal streaming step.  Usage: gpurun -- bash scripts/gpu_sanitize.sh
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 \
      python -c "import __graft_entry__ as g; g.smoke(); g.smoke_stream()" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
