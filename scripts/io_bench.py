"""Micro-benchmarks of the kernels either side of the hot path:
frame ingest (HBM GB/s against the measured copy peak), uint8 vs float backbone entry, TAP-Vid
counters, and the bulk driver against the per-frame online loop.  One JSON line per item."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tapnet_b200 import synth  # noqa: E402  (seeded weights)
from tapnet_b200 import bulk, live, metrics, tapir_model  # noqa: E402


def timeit(fn, n=20, warm=3):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / n


def main():
  peak = 6572.5
  pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
  if os.path.exists(pk):
    peak = json.load(open(pk)).get('hbm_gbs', peak)
  dev = torch.device('cuda')
  g = torch.Generator().manual_seed(0)
  for name, (T, H, W), window, res in (
      ('ingest 64x1080x1920 -> crop 1080^2 -> 1024^2', (64, 1080, 1920), (0, 420, 1080, 1080), (1024, 1024)),
      ('ingest 256x480x640 -> crop 480^2 -> 256^2', (256, 480, 640), (0, 80, 480, 480), (256, 256)),
      ('preprocess 256x256x256 (no resize)', (256, 256, 256), None, None)):
    frames = torch.randint(0, 256, (T, H, W, 3), generator=g, dtype=torch.uint8).to(dev)
    ms = timeit(lambda: live.ingest_frames(frames, window, res))
    ch, cw = (H, W) if window is None else window[2:]
    oh, ow = (ch, cw) if res is None else res
    byts = T * ch * cw * 3 + T * oh * ow * 12
    print(json.dumps(dict(item=name, ms=round(ms, 4), algorithmic_GB=round(byts / 1e9, 4),
                          GBps=round(byts / ms / 1e6, 1), frac_of_hbm_peak=round(byts / ms / 1e6 / peak, 3))))
    del frames

  model = tapir_model.TAPIR(pyramid_level=1, use_casual_conv=True)
  model.load_state_dict(synth.make_state_dict(0))
  model = model.to(dev).eval()
  frames = torch.randint(0, 256, (1, 48, 256, 256, 3), generator=g, dtype=torch.uint8).to(dev)
  video = live.preprocess_frames(frames)
  ms_u8 = timeit(lambda: model.get_feature_grids(frames, False), n=5)
  ms_f = timeit(lambda: model.get_feature_grids(live.preprocess_frames(frames), False), n=5)
  print(json.dumps(dict(item='get_feature_grids 48x256^2: uint8 fused vs preprocess+float',
                        ms_uint8=round(ms_u8, 3), ms_float=round(ms_f, 3))))
  del video

  B, N, T = 30, 1024, 250   # a TAP-Vid-DAVIS-sized evaluation batch
  rng = np.random.default_rng(0)
  gt = torch.from_numpy(rng.uniform(0, 256, (B, N, T, 2)).astype(np.float32)).to(dev)
  pred = gt + torch.randn_like(gt) * 3
  go = torch.from_numpy(rng.uniform(size=(B, N, T)) < 0.3).to(dev)
  occ, expd = torch.randn(B, N, T, device=dev), torch.randn(B, N, T, device=dev)
  qp = torch.zeros(B, N, 3, device=dev)
  ms = timeit(lambda: metrics.compute_tapvid_metrics(qp, go, gt, None, pred, 'first', pred_logits=(occ, expd)))
  byts = B * N * T * (1 + 8 + 8 + 8)
  print(json.dumps(dict(item=f'tapvid metrics fused with visibility, {B}x{N}x{T}', ms=round(ms, 4),
                        GBps=round(byts / ms / 1e6, 1))))
  t0 = time.time()
  with np.errstate(all='ignore'):
    from oracle import frames_io_oracle as io_oracle
    vis = io_oracle.postprocess_occlusions(occ.cpu(), expd.cpu())
    io_oracle.compute_tapvid_metrics(qp.cpu().numpy(), go.cpu().numpy(), gt.cpu().numpy(),
                                     (~vis).numpy(), pred.cpu().numpy(), 'first')
  print(json.dumps(dict(item='same on the host (numpy restatement of the reference)',
                        ms=round((time.time() - t0) * 1e3, 1))))

  # bulk driver: 2 videos x 48 frames, 2048 points; chunked causal vs per-frame online steps
  vids = {i: torch.randint(0, 256, (48, 256, 256, 3), generator=g, dtype=torch.uint8) for i in range(2)}
  small = {i: v[:4] for i, v in vids.items()}
  bulk.track_many_points(small, [0, 1], model, frame_stride=1, points_per_frame=32,
                         point_batch_size=2048, frames_per_step=24)   # warm-up: workspaces, packing
  for fps in (24, 1):
    torch.cuda.synchronize()
    t0 = time.time()
    r = bulk.track_many_points(vids, [0, 1], model, frame_stride=1, points_per_frame=32,
                               point_batch_size=2048, frames_per_step=fps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    npts = r['separation_tracks'][0].shape[0]
    print(json.dumps(dict(item=f'track_many_points 2x48x256^2, {npts} points, frames_per_step={fps}',
                          seconds=round(dt, 3), point_frames_per_s=round(npts * 96 / dt, 1))))


if __name__ == '__main__':
  main()
