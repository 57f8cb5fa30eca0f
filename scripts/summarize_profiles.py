"""Turns gpurun_out/ ncu captures + reports into the tracked summaries under profiles/.

Usage: python scripts/summarize_profiles.py r01
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
os.makedirs(PROF, exist_ok=True)

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
]


def raw(rep):
  p = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True)
  rows = list(csv.reader(p.stdout.splitlines()))
  if len(rows) < 3:
    return []
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}
  out = []
  for r in rows[2:]:
    d = {'kernel': re.sub(r'\(.*', '', r[idx['Kernel Name']]).replace('void ', '').replace('unnamed>::', '')}
    for m, k in WANT:
      if m in idx:
        d[k] = r[idx[m]] + ' ' + units[idx[m]]
    out.append(d)
  return out


lines = [f'# ncu summaries ({tag})', '',
         'Captured with `ncu --set full --clock-control none --import-source on` on one B200 under '
         '`gpurun` (scripts/gpu_profile.sh), bench workload (256x256x48, 256 queries).  Per-launch '
         'values are cold-cache and serialised: compare shares, not absolutes.', '']
traffic = {}
for f in sorted(os.listdir(OUT)):
  if not f.endswith('.ncu-rep'):
    continue
  recs = raw(os.path.join(OUT, f))
  lines.append(f'## {f}')
  lines.append('')
  for d in recs:
    lines.append('* ' + ', '.join(f'{k}={v}' for k, v in d.items()))
  lines.append('')

def _bytes(s):
  v, u = s.split()
  v = float(v.replace(',', ''))
  return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)

def traffic_of(rep, which=None):
  recs = raw(os.path.join(OUT, rep))
  vals = [_bytes(d['dram_read']) + _bytes(d['dram_write']) for d in recs if 'dram_read' in d]
  if not vals:
    return None
  return sum(vals) / len(vals) if which is None else vals[which]

for name, rep, which in (('backbone.extra_conv', 'prof_gemm_extra.ncu-rep', None),
                         ('mixer.up', 'prof_gemm_mixer.ncu-rep', 0),
                         ('mixer.down', 'prof_gemm_mixer.ncu-rep', 1),
                         ('mixer.dw', 'prof_mixer_dw.ncu-rep', None),
                         ('local_corr', 'prof_local_corr.ncu-rep', None),
                         ('cost_volume.head', 'prof_head.ncu-rep', None),
                         ('cost_volume.gemm', 'prof_cost_volume_gemm.ncu-rep', None),
                         ('backbone.instnorm_apply', 'prof_instnorm_apply.ncu-rep', None)):
  if os.path.exists(os.path.join(OUT, rep)):
    t = traffic_of(rep, which)
    if t:
      traffic[name] = round(t)
with open(os.path.join(PROF, 'roofline_traffic.json'), 'w') as fh:
  json.dump(traffic, fh, indent=1)

# launch list -> per-kernel shares
lp = os.path.join(OUT, 'launches.csv')
if os.path.exists(lp):
  with open(lp) as fh:
    rows = list(csv.reader(l for l in fh if not l.startswith('==')))
  hdr = rows[0]
  ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
  recs = [(re.sub(r'\(.*', '', r[ki]).split('::')[-1], float(r[vi].replace(',', ''))) for r in rows[1:] if len(r) > vi]
  with open(os.path.join(PROF, f'{tag}_launches.csv'), 'w') as fh:
    fh.write('index,kernel,duration_ns\n')
    for i, (k, v) in enumerate(recs):
      fh.write(f'{i},{k},{v:.0f}\n')
  half = recs[len(recs) // 2:]
  agg = collections.defaultdict(lambda: [0, 0.0])
  for k, v in half:
    agg[k][0] += 1
    agg[k][1] += v
  tot = sum(v[1] for v in agg.values())
  lines += [f'## launch list (second half of gpurun_out/launches.csv = one step; {len(recs)} launches captured)', '',
            '| kernel | launches | total ms | share | avg us |', '|---|---|---|---|---|']
  for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f'| {k[:70]} | {n} | {t / 1e6:.3f} | {t / tot:.1%} | {t / n / 1e3:.1f} |')
  lines.append(f'| total | | {tot / 1e6:.3f} | | |')
with open(os.path.join(PROF, f'{tag}_ncu_summary.md'), 'w') as fh:
  fh.write('\n'.join(lines) + '\n')

rp = os.path.join(OUT, 'test_report.json')
if os.path.exists(rp):
  rep = json.load(open(rp))
  with open(os.path.join(PROF, f'{tag}_parity.md'), 'w') as fh:
    fh.write(f'# GPU parity report ({tag}) - max abs errors vs oracle / golden (tests -m gpu)\n\n')
    for k, v in sorted(rep.items()):
      fh.write(f'* `{k}`: ' + ', '.join(f'{a}={b:.3g}' if isinstance(b, float) else f'{a}={b}' for a, b in v.items()) + '\n')
print('wrote', os.listdir(PROF))
