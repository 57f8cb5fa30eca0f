#!/bin/bash
# Multi-GPU pass (gpurun --gpus 8): 2-rank equality check, then bench.py (headline c2 + sub-records
# c4 strong scaling and c5) at 8, 4 and 2 GPUs, exactly as the driver launches it.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_properties_gpu.py -q -m gpu -k sharded --timeout 800 -p no:cacheprovider \
    > gpurun_out/multi_gpu_check.log 2>&1
echo "multi_gpu_check rc=$? $(tail -n 2 gpurun_out/multi_gpu_check.log | tr '\n' ' ')"
run() {  # n, output, extra args...
  local n=$1 out=$2; shift 2
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
      --master-port $((29600 + n)) bench.py --gpus "$n" --steps 10 --warmup 3 "$@" \
      > "gpurun_out/$out" 2> "gpurun_out/$out.err"
  echo "$out rc=$? $(python - "gpurun_out/$out" <<'P'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
  s = d.get('sub_records') or {}
  print('c2', d['value'], d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'),
        '| c4', (s.get('c4_strong') or {}).get('value'), (s.get('c4_strong') or {}).get('ms_per_step'),
        '| c5', (s.get('c5_hires') or {}).get('value'), (s.get('c5_hires') or {}).get('ms_per_step'))
except Exception as e:
  print('parse error', e)
P
)"
}
run 8 scale_8gpu.json
run 4 scale_4gpu.json
run 2 scale_2gpu.json
tail -n 5 gpurun_out/scale_8gpu.json.err
