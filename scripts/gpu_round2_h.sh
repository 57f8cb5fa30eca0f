#!/bin/bash
# A/B: 2-SM kernel for the N = 64 layers, weight-stationary mode
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -n 3 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
grep -E "^FAILED" gpurun_out/pytest_gpu.log | head -5
for cfg in "TAPIR_B200_GEMM_2SM_N64=1" "TAPIR_B200_GEMM_2SM_N64=1 TAPIR_B200_GEMM_BRES=1"; do
  env $cfg timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_stages_gpu.py tests/test_end_to_end_gpu.py -q -m gpu -k "not simt" --timeout 600 -p no:cacheprovider > gpurun_out/pytest_n64.log 2>&1
  echo "[$cfg] pytest rc=$? $(tail -n 2 gpurun_out/pytest_n64.log | tr '\n' ' ')"
done
i=0
for cfg in "X=1" "TAPIR_B200_GEMM_2SM_N64=1" "TAPIR_B200_GEMM_2SM_N64=1 TAPIR_B200_GEMM_BRES=1" "X=1" "TAPIR_B200_GEMM_2SM_N64=1" "TAPIR_B200_GEMM_2SM_N64=1 TAPIR_B200_GEMM_BRES=1"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --no-sub --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_n64_$i.json 2>/dev/null
  echo "[$cfg] $(python -c "
import json;d=json.load(open('gpurun_out/bench_n64_$i.json'));kb=d['kernel_breakdown']
print(d['ms_per_step'], d['clocks']['sm_mhz'], {k:kb[k]['ms_per_step'] for k in ('backbone.extra_conv','backbone.conv','backbone.proj','mixer.up','mixer.down')})")"
done
