#!/bin/bash
# Round-2 GPU pass C: suite, bench, racecheck, launch list, ncu --set full captures of the final kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -n 3 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_c.json'))
print('c2', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'e2e_float', (d.get('e2e_float_frames') or {}).get('ms_per_step'))
for k,r in (d.get('sub_records') or {}).items():
  print(k, r.get('ms_per_step', r.get('ms_per_frame')), (r.get('e2e') or {}).get('ms_per_step', (r.get('e2e') or {}).get('ms_per_frame')))
kb=d['kernel_breakdown']
print({k:v['ms_per_step'] for k,v in kb.items() if v['ms_per_step']>0.3})
P
TAPIR_B200_GEMM_TAIL=0 timeout 600 python bench.py --no-sub --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_c_notail.json 2>/dev/null
echo "no-tail: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c_notail.json | head -1)"
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); g.smoke_stream()" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer_racecheck.log | tail -1)"
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py --warm 1 --steps 1 > gpurun_out/launches.log 2>&1
echo "ncu step rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_stream.csv \
    python scripts/profile_stream.py --warm 2 --steps 1 > gpurun_out/launches_stream.log 2>&1
echo "ncu stream rc=$?"
bash scripts/gpu_profile.sh
