#!/bin/bash
# Round-2 GPU pass D: final build - suite + bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -n 3 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_d.json'))
print('c2', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'e2e_float', (d.get('e2e_float_frames') or {}).get('ms_per_step'), d['clocks'])
for k,r in (d.get('sub_records') or {}).items():
  print(k, r.get('ms_per_step', r.get('ms_per_frame')), (r.get('e2e') or {}).get('ms_per_step', (r.get('e2e') or {}).get('ms_per_frame')))
kb=d['kernel_breakdown']
print({k:v['ms_per_step'] for k,v in kb.items() if v['ms_per_step']>0.03})
print(d['roofline_named']['cost_volume']['issued_mma_frac_of_sustained_peak'], d['roofline_named']['cost_volume']['avg_launch_ms'])
kb=d['sub_records']['c3_stream']['kernel_breakdown']
print({k:(v['ms_per_step'],v['launches_per_step']) for k,v in kb.items() if v['ms_per_step']>0.04})
P
timeout 600 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_d_ref.json 2>/dev/null
echo "ref: $(python -c "import json;d=json.load(open('gpurun_out/bench_d_ref.json'));print(d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['spread'])")"
