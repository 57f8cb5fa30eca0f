#!/bin/bash
# Round-2 GPU pass A: whole -m gpu suite, default bench (with sub-records), dw-pipe A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -n 3 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench rc=$? $(head -c 300 gpurun_out/bench_a.json)"
for mode in 0 1 2; do
  TAPIR_B200_DW_PIPE=$mode timeout 600 python bench.py --workload c4 --no-sub --no-cpu --steps 4 --warmup 3 \
      > gpurun_out/bench_c4_dwpipe$mode.json 2> gpurun_out/bench_c4_dwpipe$mode.err
  echo "c4 dw_pipe=$mode rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c4_dwpipe$mode.json | head -1)"
done
TAPIR_B200_DW_PIPE=2 timeout 600 python bench.py --no-sub --no-cpu --steps 10 --warmup 3 \
    > gpurun_out/bench_c2_dwpipe2.json 2> gpurun_out/bench_c2_dwpipe2.err
echo "c2 dw_pipe=2 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c2_dwpipe2.json | head -1)"
