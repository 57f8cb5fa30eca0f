#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print('c2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'e2e_float', (d.get('e2e_float_frames') or {}).get('ms_per_step'), d['clocks'])
for k,r in (d.get('sub_records') or {}).items():
  print(k, r.get('value'), r.get('ms_per_step', r.get('ms_per_frame')), (r.get('e2e') or {}).get('ms_per_step', (r.get('e2e') or {}).get('ms_per_frame')), r['clocks']['sm_mhz'])
P
