"""CPU: host logic of the bulk multi-video driver (tapnet_b200/bulk.py)."""
import numpy as np
import pytest
import torch

from tapnet_b200 import bulk
from tapnet_b200.tapir_model import QueryFeatures


def test_sampling_follows_reference_stream():
  """np.random.seed(42) + one uniform(0,1,[P,3]) draw per selected frame, videos in order
  (tapir_clustering.py:1046,1063-1073)."""
  shapes = [(5, 64, 80, 3), (3, 64, 80, 3)]
  got = bulk.sample_query_points(shapes, 2, 4, (0.1, 0.2, 0.9, 0.7))
  np.random.seed(42)
  want = []
  for sv_idx, shp in enumerate(shapes):
    for i in range(0, shp[0], 2):
      u = np.random.uniform(0.0, 1.0, [4, 3])
      want.append((sv_idx, i, u * np.array([0.0, shp[1] * (0.7 - 0.2), shp[2] * (0.9 - 0.1)])[None]
                   + np.array([0.0, shp[1] * 0.2, shp[2] * 0.1])[None]))
  assert [(a, b) for a, b, _ in got] == [(a, b) for a, b, _ in want]
  for (_, _, p), (_, _, q) in zip(got, want):
    np.testing.assert_array_equal(p, q)
    assert (p[:, 0] == 0).all() and (p[:, 1] >= 64 * 0.2).all() and (p[:, 2] <= 80 * 0.9).all()


def test_shard_batches_covers_everything_once():
  for n in (0, 1, 5, 8):
    for world in (1, 2, 3, 8):
      seen = sorted(b for r in range(world) for b in bulk.shard_batches(n, r, world))
      assert seen == list(range(n))


def test_query_features_join_and_count():
  def qf(n, v):
    return QueryFeatures((torch.full((1, n, 256), v), torch.full((1, n, 256), v + 0.5)),
                         (torch.full((1, n, 128), v), torch.full((1, n, 128), v + 0.5)),
                         ((256, 256), (256, 256)))
  j = bulk.query_features_join([qf(2, 1.0), qf(3, 2.0)])
  assert bulk.query_features_count(j) == 5
  assert j.lowres[0].shape == (1, 5, 256) and j.hires[1].shape == (1, 5, 128)
  assert j.lowres[0][0, :, 0].tolist() == [1, 1, 2, 2, 2]
  assert j.resolutions == ((256, 256), (256, 256))


def test_predictions_to_tracks_visibility():
  p = dict(tracks=torch.zeros(1, 3, 1, 2), occlusion=torch.full((1, 3, 1), -20.0),
           expected_dist=torch.full((1, 3, 1), -20.0))
  t, v = bulk.predictions_to_tracks_visibility(p)
  assert t.shape == (3, 2) and v.shape == (3,) and (v > 0.99).all()
  t, v = bulk.predictions_to_tracks_visibility(p, single_step=False)
  assert t.shape == (3, 1, 2) and v.shape == (3, 1)


def test_rejects_cpu_and_bad_arguments():
  from tapnet_b200 import tapir_model
  m = tapir_model.TAPIR(use_casual_conv=False)
  vids = {'a': np.zeros((2, 64, 64, 3), np.uint8)}
  with pytest.raises(ValueError):
    bulk.track_many_points(vids, ['a'], m)              # needs the causal model
  m = tapir_model.TAPIR(use_casual_conv=True)
  with pytest.raises(RuntimeError):
    bulk.track_many_points(vids, ['a'], m)              # CPU module: no fallback
  with pytest.raises(ValueError):
    bulk.track_many_points(vids, ['a'], m, points_per_frame=3, point_batch_size=8)
