#!/bin/bash
# Round-2 GPU pass B: suite, bench, sanitizer, launch lists (offline step + streaming step).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -n 3 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "bench rc=$? $(head -c 200 gpurun_out/bench_b.json)"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_b.json'))
print('c2', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'e2e_float', (d.get('e2e_float_frames') or {}).get('ms_per_step'))
for k,r in (d.get('sub_records') or {}).items():
  print(k, r.get('ms_per_step', r.get('ms_per_frame')), (r.get('e2e') or {}).get('ms_per_step', (r.get('e2e') or {}).get('ms_per_frame')))
P
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_stream.csv \
    python scripts/profile_stream.py --warm 2 --steps 1 > gpurun_out/launches_stream.log 2>&1
echo "ncu stream rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py --warm 1 --steps 1 > gpurun_out/launches.log 2>&1
echo "ncu step rc=$?"
bash scripts/gpu_sanitize.sh
