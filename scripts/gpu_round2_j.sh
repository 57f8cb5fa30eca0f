#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for m in 1 2 3; do
  TAPIR_B200_CONV_HALO=$m timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "conv and tc" --timeout 200 -p no:cacheprovider > gpurun_out/pytest_halo$m.log 2>&1
  echo "halo mode $m rc=$? $(grep -E 'report\] conv_impl0_2x(32x32|20x24)' gpurun_out/pytest_halo$m.log | sed 's/\[report\] //' | tr '\n' ' ')"
done
