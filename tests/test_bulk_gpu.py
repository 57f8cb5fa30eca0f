"""-m gpu: bulk multi-video tracking (chunked causal steps on the GPU) against the CPU oracle
that runs the reference's per-frame online loop."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import bulk_oracle, synth  # noqa: E402
from oracle import tapir_oracle as O  # noqa: E402
from tapnet_b200 import bulk  # noqa: E402
from tests import gpu_util as U  # noqa: E402
from tests.test_stages_gpu import get_model  # noqa: E402


def _videos():
  rng = np.random.default_rng(0)
  base = rng.integers(0, 256, (64 + 8, 80 + 8, 3), dtype=np.uint8)
  def clip(n, dy):  # a panning crop of one random texture: frames are related, not noise
    return np.stack([base[i * dy:i * dy + 64, i:i + 80] for i in range(n)])
  return {'ep0': clip(3, 1), 'ep1': clip(5, 0)}


@pytest.mark.parametrize('frames_per_step', [2, 24])
def test_track_many_points_matches_online_oracle(frames_per_step):
  model, sd, _ = get_model(causal=True)
  cfg = O.Config(use_casual_conv=True)
  vids = _videos()
  ids = ['ep0', 'ep1']
  kw = dict(frame_stride=2, points_per_frame=4, point_batch_size=12)   # 5 frames -> 2 batches, 4 pads
  ref = bulk_oracle.track_many_points(sd, cfg, vids, ids, **kw)
  out = bulk.track_many_points(vids, ids, model, frames_per_step=frames_per_step, **kw)
  assert out['demo_episode_ids'] == ids
  assert out['video_shape'] == ref['video_shape']
  for a, b in zip(out['query_points'], ref['query_points']):
    np.testing.assert_array_equal(a, b)
  e_q = max(float(np.abs(a - b.numpy()).max()) for a, b in
            zip(out['query_features'].lowres + out['query_features'].hires,
                ref['query_features'].lowres + ref['query_features'].hires))
  e_t, flips, total = 0.0, 0, 0
  for k in ids:
    assert out['separation_tracks'][k].shape == ref['separation_tracks'][k].shape == \
        (20, vids[k].shape[0], 2)
    e_t = max(e_t, float(np.abs(out['separation_tracks'][k] - ref['separation_tracks'][k]).max()))
    decided = np.abs(ref['separation_visibility_score'][k] - 0.5) > 1e-4
    flips += int((out['separation_visibility'][k] != ref['separation_visibility'][k])[decided].sum())
    total += decided.size
  U.record(f'bulk_fps{frames_per_step}', tracks_px=e_t, qfeat=e_q, vis_flips=flips, vis_total=total)
  assert e_q <= 1e-4
  assert e_t <= 1e-3          # north-star budget for tracks
  assert flips == 0


def test_track_many_points_step_size_invariance():
  """Chunked causal execution: the result must not depend on frames_per_step beyond fp noise."""
  model, _, _ = get_model(causal=True)
  vids = _videos()
  kw = dict(frame_stride=2, points_per_frame=4, point_batch_size=8)
  a = bulk.track_many_points(vids, ['ep1', 'ep0'], model, frames_per_step=1, **kw)
  b = bulk.track_many_points(vids, ['ep1', 'ep0'], model, frames_per_step=3, **kw)
  for k in vids:
    assert float(np.abs(a['separation_tracks'][k] - b['separation_tracks'][k]).max()) <= 1e-4
