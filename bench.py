#!/usr/bin/env python
"""Benchmark of the TAPIR hot path (driver contract: see the task brief / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full inference pass  get_feature_grids + get_query_features +
estimate_trajectories (+ the per-level mean of forward)  over one synthetic clip.
Workload (BASELINE.json configs[1]): 256x256x48 video, 256 query points per GPU; with N GPUs the
job is ONE clip tracked for 256*N query points (weak scaling in queries): backbone frames are
sharded, one NCCL all-gather of the feature grids, queries sharded, no collective in the
refinement loop.  metric = query-points x frames / second, whole job.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_FRAMES, RES, Q_PER_GPU = 48, 256, 256
METRIC = 'query-points x frames / sec (TAPIR inference, 256x256x48)'
UNIT = 'point-frames/s'
SCALING, CONFIG_NAME = 'weak', 'BASELINE.json configs[1]'


def select_workload(name):
  """c2 (default, the driver's contract): 256x256x48, 256 queries per GPU, weak scaling.
  c4 (BASELINE.json configs[3], the north star's scaling target): 256x256x96, 4096 queries in
  total shared by the ranks, strong scaling."""
  global T_FRAMES, Q_PER_GPU, METRIC, SCALING, CONFIG_NAME
  if name == 'c4':
    world = int(os.environ.get('WORLD_SIZE', '1'))
    T_FRAMES, Q_PER_GPU = 96, 4096 // world
    METRIC = 'query-points x frames / sec (TAPIR inference, 256x256x96, 4096 queries)'
    SCALING, CONFIG_NAME = 'strong', 'BASELINE.json configs[3]'


def _peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as fh:
      d = json.load(fh)
    return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'],
                source='measured (MEASURED_PEAKS.json)')
  return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback (B200_PROFILING.md)')


# bf16 MMAs issued per fp32-equivalent product term under each precision policy
_MMA_TERMS = {'bf16': 1, 'bf16x3': 3, 'bf16x6': 6}


class ClockSampler:
  """nvidia-smi clocks / throttle reasons; started before warm-up, samples are kept only if
  their timestamp falls inside the timed region (mark_begin .. mark_end)."""
  Q = ('timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index = index
    self.proc = None
    self.lines = []
    self.t0 = self.t1 = None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
           '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append((time.time(), line.strip()))

  def mark_begin(self):
    self.t0 = time.time()

  def mark_end(self):
    self.t1 = time.time()

  def stop(self):
    if self.proc is None:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
    time.sleep(0.05)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def parse(window):
      sm, mx, reasons = [], None, set()
      for ts, ln in self.lines:
        if window and not (self.t0 - 0.02 <= ts <= self.t1 + 0.05):
          continue
        parts = [p.strip() for p in ln.split(',')]
        if len(parts) < 7:
          continue
        try:
          sm.append(float(parts[1]))
          mx = float(parts[2])
        except ValueError:
          continue
        for n, v in zip(names, parts[3:7]):
          if v.lower().startswith('active'):
            reasons.add(n)
      return sm, mx, reasons

    sm, mx, reasons = parse(True)
    where = 'timed region'
    if not sm:  # region shorter than the sampling period: fall back to the whole run under load
      sm, mx, reasons = parse(False)
      sm = [v for v in sm if mx and v > 0.5 * mx] or sm
      where = 'whole run (timed region shorter than the sampling period)'
    return dict(sm_mhz=(statistics.median(sm) if sm else None), sm_max_mhz=mx,
                reasons=sorted(reasons), samples=len(sm), window=where)


def build_inputs(world):
  from tapnet_b200 import synth  # seeded synthetic inputs
  video = synth.make_video(T_FRAMES, RES, RES, seed=1)
  queries = synth.make_queries(Q_PER_GPU * world, T_FRAMES, RES, RES, seed=2)
  sd = synth.make_state_dict(0)
  return sd, video, queries


# ----------------------------------------------------------------------------------- ours


def run_ours(args):
  import torch.distributed as dist
  from tapnet_b200 import _lib, tapir_model
  from tapnet_b200 import distributed as tdist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit('launch with torch.distributed.run for --gpus > 1')
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
      os.environ['NCCL_DEBUG'] = 'WARN'  # keep stdout to the single JSON line
    dist.init_process_group('nccl', device_id=dev)
  lib = _lib.load()
  sd, video_h, queries_h = build_inputs(world)
  model = tapir_model.TAPIR(pyramid_level=1, precision=args.precision)
  model.load_state_dict(sd)
  model = model.to(dev).eval()
  video_pin, queries_pin = video_h.pin_memory(), queries_h.pin_memory()
  video_d, queries_d = video_pin.to(dev), queries_pin.to(dev)
  N = queries_h.shape[1]
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
  out_pin = {k: torch.empty(s, dtype=torch.float32).pin_memory() for k, s in
             (('tracks', (1, N, T_FRAMES, 2)), ('occlusion', (1, N, T_FRAMES)),
              ('expected_dist', (1, N, T_FRAMES)))}

  def step_device():
    if world > 1:
      return tdist.sharded_forward(model, video_d, queries_d, gather_outputs=False)
    return model(video_d, queries_d)

  def step_e2e():
    v = video_pin.to(dev, non_blocking=True)
    q = queries_pin.to(dev, non_blocking=True)
    if world > 1:
      out = tdist.sharded_forward(model, v, q, gather_outputs=True)
    else:
      out = model(v, q)
    if rank == 0:
      for k in out_pin:
        out_pin[k].copy_(out[k], non_blocking=True)
    torch.cuda.current_stream().synchronize()

  # the reference's callers hold uint8 frames and normalise on the device
  # (pytorch_live_demo.py:30-41,139-141): same clip as raw frames, normalisation fused into the
  # stem conv, a quarter of the PCIe bytes
  frames_u8_pin = ((video_h + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).pin_memory()

  def step_e2e_u8():
    v = frames_u8_pin.to(dev, non_blocking=True)
    q = queries_pin.to(dev, non_blocking=True)
    if world > 1:
      out = tdist.sharded_forward(model, v, q, gather_outputs=True)
    else:
      out = model(v, q)
    if rank == 0:
      for k in out_pin:
        out_pin[k].copy_(out[k], non_blocking=True)
    torch.cuda.current_stream().synchronize()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    """K steps inside one barrier+sync bracket; per-step CUDA events; L2 flushed between steps."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(steps)]
    barrier()
    for a, b in evs:
      flush.zero_()
      a.record()
      fn()
      b.record()
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / steps

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  for _ in range(max(args.warmup, 3)):
    step_device()
  barrier()
  launches0 = lib.tapir_launch_count()
  sampler.mark_begin()
  ms_step = timed(step_device, args.steps)
  sampler.mark_end()
  launches = (lib.tapir_launch_count() - launches0) // max(args.steps, 1)
  clocks = sampler.stop() if rank == 0 else None
  step_e2e()
  ms_e2e = timed(step_e2e, args.steps)
  step_e2e_u8()
  ms_e2e_u8 = timed(step_e2e_u8, args.steps)

  # per-kernel device time (CUDA events around every launch of this library, on its stream):
  # two extra steps after the timed region
  prof = None
  lib.tapir_profile_enable(1)
  step_device()
  step_device()
  import ctypes
  cbuf = ctypes.create_string_buffer(1 << 16)
  if lib.tapir_profile_report(cbuf, len(cbuf)) == 0:
    prof = json.loads(cbuf.value.decode())
  lib.tapir_profile_enable(0)

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  units = N * T_FRAMES
  peaks = _peaks()
  roofline = None
  breakdown = None
  if prof:
    tot = sum(v['ms'] for v in prof.values())
    breakdown = {k: dict(ms_per_step=round(v['ms'] / 2, 4), launches_per_step=v['launches'] // 2,
                         share=round(v['ms'] / tot, 4),
                         tflops=round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 2) if v['ms'] > 0 else 0,
                         gbs=round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1) if v['ms'] > 0 else 0)
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])}
    top = max(prof.items(), key=lambda kv: kv[1]['ms'])
    name, v = top
    is_gemm = v['flops'] > 0 and ('mixer.' in name or 'conv' in name or 'gemm' in name or 'proj' in name) \
        and name != 'mixer.dw'
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
      with open(tpath) as fh:
        traffic = json.load(fh).get(name)
    if is_gemm:
      ach = v['flops'] / (v['ms'] * 1e-3) / 1e12
      roofline = dict(kernel=name, bound='tensor', achieved=round(ach, 2), peak=peaks['tf_sustained'],
                      unit='TFLOP/s', frac=round(ach / peaks['tf_sustained'], 4), traffic=traffic,
                      peak_source=peaks['source'] + ', sustained (kernel timed inside a long step)',
                      launches=v['launches'] // 2, avg_launch_ms=round(v['ms'] / v['launches'], 4),
                      mma_terms=_MMA_TERMS[args.precision],
                      issued_mma_tflops=round(ach * _MMA_TERMS[args.precision], 1),
                      issued_mma_frac_of_burst_peak=round(
                          ach * _MMA_TERMS[args.precision] / peaks['tf_burst'], 4),
                      note='achieved counts ALGORITHMIC fp32-equivalent FLOPs (2*M*N*K); the kernel '
                           'issues 3 bf16 MMAs per product term (split-bf16, required by the 1e-4 '
                           'parity budget), so tensor-pipe work is 3x this figure')
    else:
      ach = v['bytes'] / (v['ms'] * 1e-3) / 1e9
      roofline = dict(kernel=name, bound='hbm', achieved=round(ach, 1), peak=peaks['hbm'], unit='GB/s',
                      frac=round(ach / peaks['hbm'], 4), traffic=traffic, peak_source=peaks['source'],
                      launches=v['launches'] // 2, avg_launch_ms=round(v['ms'] / v['launches'], 4))
  cpu_baseline = cpu_reference_sample(sd, video_h, queries_h) if world == 1 and not args.no_cpu else None
  line = dict(
      metric=METRIC, value=round(units / (ms_step * 1e-3), 1), unit=UNIT, n_gpus=world,
      steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=round(ms_step, 3),
      higher_is_better=True, scaling=SCALING, vs_baseline=None,
      dtype='bf16x3' if args.precision == 'bf16x3' else args.precision, data='synthetic',
      config=dict(workload=f'TAPIR/BootsTAPIR inference {RES}x{RES}x{T_FRAMES}, {Q_PER_GPU} query '
                           f'points per GPU ({N} total), {CONFIG_NAME}',
                  frames=T_FRAMES, resolution=RES, queries=N, refine_iterations=4,
                  parallelism=f'frame-shard backbone + all-gather + query-shard x{world}',
                  l2='256 MiB buffer written between timed steps (L2 flush)',
                  weights='seeded random init (no checkpoint reachable offline)'),
      clocks=clocks,
      e2e=dict(value=round(units / (ms_e2e * 1e-3), 1), unit=UNIT, ms_per_step=round(ms_e2e, 3),
               h2d_bytes_per_step=int(video_h.numel() * 4 + queries_h.numel() * 4),
               d2h_bytes_per_step=int(sum(t.numel() * 4 for t in out_pin.values()))),
      e2e_uint8_frames=dict(
          value=round(units / (ms_e2e_u8 * 1e-3), 1), unit=UNIT, ms_per_step=round(ms_e2e_u8, 3),
          h2d_bytes_per_step=int(video_h.numel() + queries_h.numel() * 4),
          d2h_bytes_per_step=int(sum(t.numel() * 4 for t in out_pin.values())),
          note='same call with the clip as raw uint8 frames (what the reference callers hold, '
               'pytorch_live_demo.py:30-41); preprocess_frames is fused into the stem conv'),
      gpu_launches=int(launches),
      roofline=roofline, cpu_baseline=cpu_baseline, kernel_breakdown=breakdown,
  )
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


# ----------------------------------------------------------------------------------- CPU reference


_THREADS = None


def _best_thread_count(sd, video):
  """torch's CPU kernels do not scale to very wide hosts on these small per-frame problems
  (128 threads were 40x slower than 8 on the GPU box), so give the reference its best case:
  time one backbone frame at a few thread counts and keep the fastest."""
  global _THREADS
  if _THREADS is not None:
    return _THREADS
  from oracle import tapir_oracle as O
  ncpu = os.cpu_count() or 1
  cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
  best, best_t = cands[0], float('inf')
  cfg = O.Config()
  with torch.no_grad():
    for c in cands:
      torch.set_num_threads(c)
      O.get_feature_grids(sd, cfg, video[:, :1])  # warm
      t0 = time.perf_counter()
      O.get_feature_grids(sd, cfg, video[:, :2])
      dt = time.perf_counter() - t0
      if dt < best_t:
        best, best_t = c, dt
  _THREADS = best
  return best


def cpu_reference_sample(sd, video, queries, sample_queries=32, sample_frames_backbone=8):
  """Reference algorithm (oracle port of tapnet/torch, fp32, torch CPU ops) on the host cores.

  Bounded sample of the SAME workload: the backbone is timed on `sample_frames_backbone` of the
  48 frames (frames are independent, cost is linear in frames), stage A + refinement on
  `sample_queries` of the queries over all 48 frames (queries are independent, linear).  The
  per-unit costs are scaled to the full job: value = N*T / (t_backbone_full + t_refine_full).
  """
  from oracle import tapir_oracle as O
  cores = _best_thread_count(sd, video)
  torch.set_num_threads(cores)
  cfg = O.Config()
  N = queries.shape[1]
  T = video.shape[1]
  with torch.no_grad():
    t0 = time.perf_counter()
    O.get_feature_grids(sd, cfg, video[:, :sample_frames_backbone])
    t_bb = (time.perf_counter() - t0) * (T / sample_frames_backbone)
    # features for the refinement sample: reuse random unit grids of the right shape (timing only)
    g = torch.Generator().manual_seed(0)
    lo = torch.nn.functional.normalize(torch.randn(1, T, 32, 32, 256, generator=g), dim=-1)
    hi = torch.nn.functional.normalize(torch.randn(1, T, 64, 64, 128, generator=g), dim=-1)
    grids = O.Grids((lo, lo), (hi, hi), ((RES, RES), (RES, RES)))
    qs = queries[:, :sample_queries]
    t0 = time.perf_counter()
    qf = O.get_query_features(cfg, video.shape, qs, grids)
    O.estimate_trajectories(sd, cfg, (RES, RES), grids, qf, qs, 64)
    t_rf = (time.perf_counter() - t0) * (N / sample_queries)
  total = t_bb + t_rf
  return dict(value=round(N * T / total, 1), unit=UNIT, cores=cores, kind='port',
              sample=f'backbone on {sample_frames_backbone}/{T} frames, stage A + 4 refine iterations '
                     f'on {sample_queries}/{N} queries x {T} frames; costs scaled linearly to the '
                     f'full job (est. {total:.1f} s/step)',
              note='reference JAX-CPU path cannot run (no jax in the image); this is the CPU '
                   'restatement of the reference torch path (oracle/tapir_oracle.py), '
                   f"torch {torch.__version__}, {cores} threads (best of a thread-count sweep, {os.cpu_count()} logical cores)")


def run_reference(args):
  """--impl reference: the reference's own CPU implementation of the path (oracle port; the
  reference is Python and /root/reference does not exist on the GPU box)."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
  sd, video, queries = build_inputs(world)
  vals = []
  for i in range(args.warmup + args.steps):
    r = cpu_reference_sample(sd, video, queries, sample_queries=16, sample_frames_backbone=4)
    if i >= args.warmup:
      vals.append(r)
  v = statistics.median([r['value'] for r in vals])
  N = queries.shape[1]
  base = vals[0]
  base['value'] = v
  line = dict(impl='reference', metric=METRIC, value=v, unit=UNIT, n_gpus=world, steps=args.steps,
              warmup=args.warmup, ms_per_step=round(N * T_FRAMES / v * 1e3, 1), higher_is_better=True,
              scaling=SCALING, vs_baseline=None, dtype='f32', data='synthetic',
              config=dict(workload=f'TAPIR/BootsTAPIR inference {RES}x{RES}x{T_FRAMES}, {Q_PER_GPU} query '
                                   f'points per GPU ({N} total), {CONFIG_NAME}',
                          frames=T_FRAMES, resolution=RES, queries=N, refine_iterations=4),
              cpu_baseline=base,
              e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--precision', default='bf16x3', choices=['bf16', 'bf16x3', 'bf16x6'])
  ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
  ap.add_argument('--workload', default='c2', choices=['c2', 'c4'],
                  help='c2 = driver contract (default); c4 = 4096 queries x 96 frames, strong scaling')
  args = ap.parse_args()
  select_workload(args.workload)
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
