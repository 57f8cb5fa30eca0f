#!/bin/bash
# halo-patch conv: parity, then A/B on the headline
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
TAPIR_B200_CONV_HALO=1 timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "conv and tc or stat" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_halo.log 2>&1
echo "halo gemm tests rc=$? $(tail -n 2 gpurun_out/pytest_halo.log | tr '\n' ' ')"
grep -E "^FAILED|Error|error" gpurun_out/pytest_halo.log | head -8
TAPIR_B200_CONV_HALO=1 timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_end_to_end_gpu.py -q -m gpu -k "backbone or golden" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_halo2.log 2>&1
echo "halo backbone/e2e rc=$? $(tail -n 2 gpurun_out/pytest_halo2.log | tr '\n' ' ')"
grep -E "^FAILED" gpurun_out/pytest_halo2.log | head -8
i=0
for cfg in "X=1" "TAPIR_B200_CONV_HALO=1" "X=1" "TAPIR_B200_CONV_HALO=1"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --no-sub --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_halo_$i.json 2>/dev/null
  echo "[$cfg] $(python -c "
import json;d=json.load(open('gpurun_out/bench_halo_$i.json'));kb=d['kernel_breakdown']
print(d['ms_per_step'], d['clocks']['sm_mhz'], {k:kb[k]['ms_per_step'] for k in ('backbone.extra_conv','backbone.conv','mixer.up')})")"
done
