"""CPU: the restatement of the steps either side of the path (oracle/frames_io_oracle.py) against
fixtures produced by the reference's own functions (oracle/make_golden_io.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import frames_io_oracle as io_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def ingest_golden():
  return np.load(os.path.join(GOLDEN, 'io_ingest.npz'))


@pytest.fixture(scope='module')
def tapvid_golden():
  return np.load(os.path.join(GOLDEN, 'io_tapvid.npz'))


@pytest.mark.parametrize('tag', ['land', 'port', 'square', 'up'])
def test_ingest_matches_reference(ingest_golden, tag):
  g = ingest_golden
  frames = torch.from_numpy(g[f'{tag}_frames'])
  window = io_oracle.center_square_window(frames.shape[1], frames.shape[2])
  assert (frames.shape[0], window[2], window[3], 3) == tuple(g[f'{tag}_cropped_shape'])
  pre = io_oracle.ingest(frames, window, None)
  np.testing.assert_array_equal(pre.numpy(), g[f'{tag}_preprocessed'])
  out = io_oracle.ingest(frames, window, tuple(g[f'{tag}_resolution']))
  np.testing.assert_allclose(out.numpy(), g[f'{tag}_out'], rtol=0, atol=1e-6)


def test_center_square_window_quirk():
  assert io_oracle.center_square_window(240, 320) == (0, 40, 240, 240)
  assert io_oracle.center_square_window(320, 240) == (40, 0, 240, 240)
  assert io_oracle.center_square_window(36, 53) == (0, 8, 36, 37)   # odd difference: not square
  with pytest.raises(ValueError):
    io_oracle.center_square_window(240, 241)                         # reference slices [0:-0]


def test_postprocess_matches_reference(ingest_golden):
  g = ingest_golden
  vis = io_oracle.postprocess_occlusions(torch.from_numpy(g['occ_logits']),
                                         torch.from_numpy(g['expd_logits']))
  np.testing.assert_array_equal(vis.numpy(), g['visible'])


@pytest.mark.parametrize('mode', ['first', 'strided'])
@pytest.mark.parametrize('trackwise', [False, True])
def test_tapvid_metrics_match_reference(tapvid_golden, mode, trackwise):
  g = tapvid_golden
  m = io_oracle.compute_tapvid_metrics(g['query_points'], g['gt_occluded'], g['gt_tracks'],
                                       g['pred_occluded'], g['pred_tracks'], mode,
                                       get_trackwise_metrics=trackwise)
  prefix = f'{mode}_{"track" if trackwise else "video"}_'
  keys = [k[len(prefix):] for k in g.files if k.startswith(prefix)]
  assert sorted(keys) == sorted(m.keys()) and len(keys) == 13
  for k in keys:
    np.testing.assert_array_equal(m[k], g[prefix + k], err_msg=k)  # nan == nan here


def test_tapvid_bad_mode():
  with pytest.raises(ValueError):
    io_oracle.compute_tapvid_metrics(np.zeros((1, 1, 3)), np.zeros((1, 1, 2), bool),
                                     np.zeros((1, 1, 2, 2)), np.zeros((1, 1, 2), bool),
                                     np.zeros((1, 1, 2, 2)), 'last')


@pytest.mark.parametrize('trackwise', [False, True])
def test_product_ratio_logic_matches_reference_from_oracle_counts(tapvid_golden, trackwise):
  """tapnet_b200.metrics.metrics_from_counts (the host half of the device metrics; runs on CPU
  tensors too) applied to the oracle's counters reproduces the reference's 13 metrics bit for bit."""
  from tapnet_b200 import metrics
  g = tapvid_golden
  for mode in ('first', 'strided'):
    counts = io_oracle.tapvid_counts(g['query_points'], g['gt_occluded'], g['gt_tracks'],
                                     g['pred_occluded'], g['pred_tracks'], mode)
    assert counts.shape == (3, 9, 18) and counts.dtype == np.int32
    m = metrics.metrics_from_counts(torch.from_numpy(counts), trackwise)
    prefix = f'{mode}_{"track" if trackwise else "video"}_'
    for k, v in m.items():
      np.testing.assert_array_equal(v.numpy(), g[prefix + k], err_msg=k)


def test_counts_are_consistent():
  """Counter identities that hold for any input: tp <= correct <= visible <= evaluated, etc."""
  rng = np.random.default_rng(8)
  B, N, T = 2, 5, 17
  qp = np.zeros((B, N, 3), np.float32)
  qp[..., 0] = rng.integers(0, T, (B, N))
  gt = rng.uniform(0, 64, (B, N, T, 2)).astype(np.float32)
  pred = (gt + rng.normal(0, 4, gt.shape)).astype(np.float32)
  go, po = rng.uniform(size=(B, N, T)) < 0.5, rng.uniform(size=(B, N, T)) < 0.5
  for mode in ('first', 'strided'):
    c = io_oracle.tapvid_counts(qp, go, gt, po, pred, mode).astype(np.int64)
    assert (c[..., 1] <= c[..., 0]).all() and (c[..., 2] <= c[..., 0]).all()
    for i in range(5):
      assert (c[..., 8 + i] <= c[..., 3 + i]).all() and (c[..., 3 + i] <= c[..., 2]).all()
      if i:
        assert (c[..., 3 + i] >= c[..., 2 + i]).all()      # wider threshold, more points inside
        assert (c[..., 13 + i] <= c[..., 12 + i]).all()    # ... and fewer false positives
    if mode == 'strided':
      assert (c[..., 0] == T - 1).all()
    else:
      assert (c[..., 0] == T - 1 - np.round(qp[..., 0]).astype(np.int64)).all()
