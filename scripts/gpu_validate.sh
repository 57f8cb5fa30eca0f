#!/bin/bash
# Full GPU validation pass: the whole -m gpu suite, then the default bench line (headline + sub-records).
# Usage: gpurun -- bash scripts/gpu_validate.sh
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -n 4 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head -10
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_e.json'))
print('c2', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'e2e_float', (d.get('e2e_float_frames') or {}).get('ms_per_step'), d['clocks'])
for k,r in (d.get('sub_records') or {}).items():
  print(k, r.get('ms_per_step', r.get('ms_per_frame')), (r.get('e2e') or {}).get('ms_per_step', (r.get('e2e') or {}).get('ms_per_frame')))
P
