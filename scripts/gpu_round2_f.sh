#!/bin/bash
# policy A/B: tile-width bias of the 2-SM kernel; SIMT cross-check tiers
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for b in 9 10 9 10; do
  TAPIR_B200_GEMM_BIAS=$b timeout 600 python bench.py --no-sub --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_bias$b.json 2>/dev/null
  echo "bias=$b $(python -c "
import json;d=json.load(open('gpurun_out/bench_bias$b.json'));kb=d['kernel_breakdown']
print(d['ms_per_step'], d['clocks']['sm_mhz'], {k:kb[k]['ms_per_step'] for k in ('backbone.extra_conv','backbone.conv','mixer.up','mixer.down','cost_volume.gemm')})")"
done
bash scripts/gpu_ci.sh stages_simt e2e_simt 2>&1 | grep -E "tier|passed|failed" | tail -6
