#!/bin/bash
# Usage: gpurun -- bash scripts/gpu_profile_one.sh <name> <kernel-regex> <skip> <count>
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -k "regex:$2" -s "$3" -c "$4" -f \
    -o "gpurun_out/prof_$1" python scripts/profile_step.py --warm 1 --steps 1 > "gpurun_out/prof_$1.log" 2>&1
echo "capture $1 rc=$?"
