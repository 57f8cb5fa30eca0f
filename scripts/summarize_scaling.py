"""profiles/r02_scaling.md from gpurun_out/scale_{2,4,8}gpu.json (+ the 1-GPU bench line)."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
one = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'profiles', f'{tag}_bench_1gpu.json')


def load(path):
  with open(path) as fh:
    return json.loads([l for l in fh if l.startswith('{')][-1])


recs = {1: load(one)}
for n in (2, 4, 8):
  p = os.path.join(ROOT, 'gpurun_out', f'scale_{n}gpu.json')
  if os.path.exists(p):
    try:
      recs[n] = load(p)
      shutil.copy(p, os.path.join(ROOT, 'profiles', f'{tag}_scale_{n}gpu.json'))
    except Exception as e:  # pylint: disable=broad-except
      print('skip', p, e)
L = [f'# Multi-GPU lines ({tag}): `bench.py --gpus N` under torchrun, one process per GPU, NCCL', '',
     'Same launch as the driver\'s (`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N '
     '--steps 10 --warmup 3`); values are whole-job point-frames/s, device-timed, max over ranks; '
     'sub-records come from the same process.', '',
     '## Headline, BASELINE config 2 (256x256x48, 256 queries PER GPU: weak scaling in queries)', '',
     '| GPUs | queries | point-frames/s | ms per step | e2e (host buffers) ms | SM MHz |', '|---|---|---|---|---|---|']
for n, d in sorted(recs.items()):
  e = (d.get('e2e') or {}).get('ms_per_step')
  L.append(f"| {n} | {d['config']['queries']} | {d['value']:.0f} | {d['ms_per_step']} | {e} | {(d.get('clocks') or {}).get('sm_mhz')} |")
L += ['', '## `c4_strong`: BASELINE config 4 (256x256x96, 4096 queries IN TOTAL shared by the ranks)', '',
      '| GPUs | point-frames/s | ms per step | speed-up vs 1 GPU | e2e ms | SM MHz |', '|---|---|---|---|---|---|']
base = None
for n, d in sorted(recs.items()):
  r = (d.get('sub_records') or {}).get('c4_strong')
  if not r or 'value' not in r:
    continue
  base = base or r['value']
  e = (r.get('e2e') or {}).get('ms_per_step')
  L.append(f"| {n} | {r['value']:.0f} | {r['ms_per_step']} | {r['value'] / base:.2f}x | {e} | {(r.get('clocks') or {}).get('sm_mhz')} |")
L += ['', '## `c5_hires`: BASELINE config 5 (1024x1024x64, 8192 queries in total, three refinement levels)', '',
      '| GPUs | point-frames/s | ms per step | speed-up vs 1 GPU | SM MHz |', '|---|---|---|---|---|']
base = None
for n, d in sorted(recs.items()):
  r = (d.get('sub_records') or {}).get('c5_hires')
  if not r or 'value' not in r:
    continue
  base = base or r['value']
  L.append(f"| {n} | {r['value']:.0f} | {r['ms_per_step']} | {r['value'] / base:.2f}x | {(r.get('clocks') or {}).get('sm_mhz')} |")
chk = os.path.join(ROOT, 'gpurun_out', 'multi_gpu_check.log')
if os.path.exists(chk):
  with open(chk) as fh:
    tail = [l.strip() for l in fh if l.strip()][-3:]
  L += ['', '## Sharded == single GPU (tests/test_properties_gpu.py::test_c4_sharded_equals_single_gpu -> tests/multi_gpu_check.py, 2 ranks)', '',
        '```'] + tail + ['```']
with open(os.path.join(ROOT, 'profiles', f'{tag}_scaling.md'), 'w') as fh:
  fh.write('\n'.join(L) + '\n')
print('\n'.join(L))
