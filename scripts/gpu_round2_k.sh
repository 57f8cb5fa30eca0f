#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -n 3 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
grep -E "^FAILED" gpurun_out/pytest_gpu.log | head -5
i=0
for cfg in "TAPIR_B200_CONV_HALO=0" "X=1" "TAPIR_B200_CONV_HALO=0" "X=1"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --no-sub --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_halo_$i.json 2>/dev/null
  echo "[$cfg] $(python -c "
import json;d=json.load(open('gpurun_out/bench_halo_$i.json'));kb=d['kernel_breakdown']
print(d['ms_per_step'], d['clocks']['sm_mhz'], {k:kb[k]['ms_per_step'] for k in ('backbone.extra_conv','backbone.conv','mixer.up')})")"
done
