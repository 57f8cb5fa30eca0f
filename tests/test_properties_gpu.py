"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there).

c2: 256x256x48, 256 queries; c3: causal streaming, 1024 queries; c5: 1024x1024 (3 refinement
levels) on a reduced clip against the oracle; c4 (multi-GPU): tests/multi_gpu_check.py.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import synth  # noqa: E402
from oracle import tapir_oracle as O  # noqa: E402
from tapnet_b200 import tapir_model  # noqa: E402
from tests import gpu_util as U  # noqa: E402
from tests.test_stages_gpu import get_model  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c2_chunk_and_permutation_invariance_and_passthrough():
  model, _, _ = get_model()
  T, N = 48, 256
  video = synth.make_video(T).cuda()
  q = synth.make_queries(N, T).cuda()
  out = model(video, q)
  # (a) different chunking + permuted queries: queries are independent, so nothing may change
  perm = torch.randperm(N, generator=torch.Generator().manual_seed(7)).cuda()
  old = tapir_model._MAX_ROWS_PER_CHUNK
  tapir_model._MAX_ROWS_PER_CHUNK = 37 * T  # ragged chunks of 37 queries
  try:
    out_p = model(video, q[:, perm])
  finally:
    tapir_model._MAX_ROWS_PER_CHUNK = old
  e_t = (out_p['tracks'] - out['tracks'][:, perm]).abs().max().item()
  e_o = (out_p['occlusion'] - out['occlusion'][:, perm]).abs().max().item()
  # (b) iteration-0 tracks pass exactly through the query point on the query frame
  # (utils.py:171-191); later iterations move it
  t0 = out['unrefined_tracks'][0][0]  # [N, T, 2]
  qi = q[0, :, 0].long()
  got = t0[torch.arange(N, device='cuda'), qi]
  e_q = (got - torch.flip(q[0, :, 1:3], dims=(-1,))).abs().max().item()
  # (c) stage A of a frame does not depend on the other frames (T-slice independence)
  half = model(video[:, :24], torch.cat([q[..., :1].clamp(max=23), q[..., 1:]], -1))
  sel = (q[0, :, 0] < 24)
  e_s = (half['unrefined_tracks'][0][0, sel] - t0[sel, :24]).abs().max().item()
  e_so = (half['unrefined_occlusion'][0][0, sel] - out['unrefined_occlusion'][0][0, sel, :24]).abs().max().item()
  U.record('c2_properties', chunk_perm_tracks=e_t, chunk_perm_occ=e_o, query_passthrough=e_q,
           tslice_tracks=e_s, tslice_occ=e_so)
  assert e_t <= 1e-4 and e_o <= 1e-5
  assert e_q == 0.0
  assert e_s <= 1e-3 and e_so <= 1e-4
  assert torch.isfinite(out['tracks']).all() and torch.isfinite(out['expected_dist']).all()


def test_c3_streaming_equals_offline_causal_1024_queries():
  """frame-by-frame with causal state == one offline call of the causal model."""
  model, _, _ = get_model(causal=True)
  T, N = 6, 1024
  video = synth.make_video(T).cuda()
  q = synth.make_queries(N, T, frame0_only=True).cuda()
  g0 = model.get_feature_grids(video[:, :1], False)
  qf = model.get_query_features(video[:, :1], False, q, g0)
  state = model.construct_initial_causal_state(N, len(qf.resolutions) - 1)
  state = [{k: v.cuda() for k, v in d.items()} for d in state]
  tr, oc = [], []
  for t in range(T):
    gr = model.get_feature_grids(video[:, t:t + 1], False)
    r = model.estimate_trajectories((256, 256), False, gr, qf, None, 64, causal_context=state,
                                    get_causal_context=True)
    state = r['causal_context']
    tr.append(r['tracks'][-1])
    oc.append(r['occlusion'][-1])
  full = model.get_feature_grids(video, False)
  off = model.estimate_trajectories((256, 256), False, full, qf, None, 64)
  e_t = (torch.cat(tr, 2) - off['tracks'][-1]).abs().max().item()
  e_o = (torch.cat(oc, 2) - off['occlusion'][-1]).abs().max().item()
  U.record('c3_streaming_vs_offline', tracks=e_t, occ=e_o)
  assert e_t <= 1e-3 and e_o <= 1e-4


def test_c5_hires_1024_against_oracle_reduced_clip():
  """1024x1024 -> refinement at 256/512/1024 (12 iterations); one frame, 8 queries."""
  model, sd, cfg = get_model()
  video = synth.make_video(1, 1024, 1024)
  q = synth.make_queries(8, 1, 1024, 1024)
  torch.set_num_threads(min(16, os.cpu_count() or 1))
  with torch.no_grad():
    ref = O.forward(sd, cfg, video, q)
  out = model(video.cuda(), q.cuda())
  assert len(out['unrefined_tracks']) == 12
  e_t = (out['tracks'].cpu() - ref['tracks']).abs().max().item()
  e_o = (out['occlusion'].cpu() - ref['occlusion']).abs().max().item()
  e_e = (out['expected_dist'].cpu() - ref['expected_dist']).abs().max().item()
  U.record('c5_1024_vs_oracle', tracks=e_t, occ=e_o, expd=e_e)
  # tracks are in 1024-pixel units here: 1e-3 px at 256 scale = 4e-3
  assert e_t <= 4e-3 and e_o <= 1e-4 and e_e <= 1e-4


def test_c3_graph_replay_tracker_equals_eager_streaming():
  """tapnet_b200.streaming.OnlineTracker (CUDA-graph replay) == eager per-frame calls, bit for bit."""
  from tapnet_b200 import streaming
  model, _, _ = get_model(causal=True)
  T, N = 5, 64
  video = synth.make_video(T).cuda()
  q = synth.make_queries(N, T, frame0_only=True).cuda()
  trk = streaming.OnlineTracker(model, 256, 256, N)
  qf = trk.init(video[0, 0], q)
  state = model.construct_initial_causal_state(N, len(qf.resolutions) - 1)
  state = [{k: v.cuda().clone() for k, v in d.items()} for d in state]
  for t in range(T):
    tracks, vis = trk.step(video[0, t])
    gr = model.get_feature_grids(video[:, t:t + 1], False)
    r = model.estimate_trajectories((256, 256), False, gr, qf, None, 64, causal_context=state,
                                    get_causal_context=True)
    state = r['causal_context']
    assert torch.equal(tracks, r['tracks'][-1]), f'frame {t}'
    occ, expd = trk.last_logits
    assert torch.equal(occ, r['occlusion'][-1]) and torch.equal(expd, r['expected_dist'][-1])
  U.record('c3_graph_tracker', frames=T, equal=1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_c4_sharded_equals_single_gpu():
  port = 29500 + os.getpid() % 1000
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', str(port),
         os.path.join(ROOT, 'tests', 'multi_gpu_check.py')]
  p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
  print(p.stdout[-2000:], p.stderr[-2000:])
  assert p.returncode == 0 and 'MULTI_GPU_OK' in p.stdout
