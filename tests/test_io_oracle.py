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
