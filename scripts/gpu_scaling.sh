#!/bin/bash
# Multi-GPU pass (gpurun --gpus 8): 2-rank equality check, then the strong-scaling workload
# (BASELINE config 4: 4096 queries x 96 frames) at 8 and 4 GPUs and the contract workload at 8.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
run() {  # n, extra args..., output
  local n=$1; shift
  local out=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
      --master-port $((29600 + n)) bench.py --gpus "$n" --steps 10 --warmup 3 --no-cpu "$@" \
      > "gpurun_out/$out" 2> "gpurun_out/$out.err"
  echo "$out rc=$? $(tail -c 100000 gpurun_out/$out | grep -o '"value": [0-9.]*' | head -1)"
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29555 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1
echo "multi_gpu_check rc=$? $(grep -E 'MULTI_GPU|errors' gpurun_out/multi_gpu_check.log | tr '\n' ' ')"
run 8 scale_c4_8gpu.json --workload c4
run 4 scale_c4_4gpu.json --workload c4
run 8 scale_c2_8gpu.json
