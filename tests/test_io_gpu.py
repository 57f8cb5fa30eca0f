"""-m gpu: the kernels either side of the hot path (frame ingest, visibility, TAP-Vid counters)
against the CPU oracle and the reference-generated fixtures, through the C ABI."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import frames_io_oracle as io_oracle  # noqa: E402
from tapnet_b200 import live, metrics  # noqa: E402
from tests import gpu_util as U  # noqa: E402
from tests.test_stages_gpu import get_model, maxerr  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('tag', ['land', 'port', 'square', 'up'])
def test_ingest_golden(tag):
  g = np.load(os.path.join(GOLDEN, 'io_ingest.npz'))
  frames = torch.from_numpy(g[f'{tag}_frames']).cuda()
  window = live.center_square_window(frames.shape[1], frames.shape[2])
  pre = live.ingest_frames(frames, window)
  np.testing.assert_array_equal(pre.cpu().numpy(), g[f'{tag}_preprocessed'])  # bit-exact
  out = live.ingest_frames(frames, window, tuple(g[f'{tag}_resolution']))
  err = float(np.abs(out.cpu().numpy() - g[f'{tag}_out']).max())
  U.record(f'ingest_golden_{tag}', max_abs_err=err)
  assert err <= 2e-6   # fp32 interpolation of values in [-1,1]; contraction order only


def test_preprocess_all_byte_values_exact():
  frames = torch.arange(256, dtype=torch.uint8).repeat(3)[: 16 * 16 * 3].reshape(1, 16, 16, 3).cuda()
  out = live.preprocess_frames(frames)
  ref = io_oracle.preprocess_frames(frames.cpu())
  assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize('hw,window,res', [((240, 320), 'center', (256, 256)),
                                           ((1080, 1920), (100, 200, 800, 1024), (512, 512)),
                                           ((64, 64), None, (256, 256)),
                                           ((480, 480), None, None)])
def test_ingest_vs_oracle(hw, window, res):
  g = torch.Generator().manual_seed(3)
  frames = torch.randint(0, 256, (2, hw[0], hw[1], 3), generator=g, dtype=torch.uint8)
  if window == 'center':
    window = live.center_square_window(*hw)
  ref = io_oracle.ingest(frames, window, res)
  out = live.ingest_frames(frames.cuda(), window, res)
  assert out.shape == ref.shape
  err = maxerr(out, ref)
  U.record(f'ingest_{hw[0]}x{hw[1]}', max_abs_err=err)
  assert err <= 2e-6


def test_ingest_errors():
  f = torch.zeros(1, 8, 8, 3, dtype=torch.uint8).cuda()
  with pytest.raises(Exception):
    live.ingest_frames(f, (0, 4, 8, 8))          # window leaves the frame
  with pytest.raises(ValueError):
    live.ingest_frames(f.float())                # not uint8
  with pytest.raises(RuntimeError):
    live.ingest_frames(f.cpu())                  # no CPU fallback


def test_backbone_u8_equals_float_path():
  """uint8 frames through the fused stem == preprocess_frames then the float path, bit for bit;
  and with a resize level (ingest kernel) too."""
  model, _, _ = get_model()
  g = torch.Generator().manual_seed(5)
  frames = torch.randint(0, 256, (1, 3, 256, 256, 3), generator=g, dtype=torch.uint8).cuda()
  video = live.preprocess_frames(frames)
  a = model.get_feature_grids(frames, False)
  b = model.get_feature_grids(video, False)
  for x, y in zip(a.lowres + a.hires, b.lowres + b.hires):
    assert torch.equal(x, y)
  small = frames[:, :, :128, :160].contiguous()   # refined at 256^2 after a resize
  a = model.get_feature_grids(small, False, refinement_resolutions=[(256, 256), (128, 160)])
  b = model.get_feature_grids(live.preprocess_frames(small), False,
                              refinement_resolutions=[(256, 256), (128, 160)])
  for x, y in zip(a.lowres + a.hires, b.lowres + b.hires):
    assert maxerr(x, y) <= 1e-5   # ingest fuses normalise + resize (same arithmetic order)


def test_postprocess_occlusions():
  g = np.load(os.path.join(GOLDEN, 'io_ingest.npz'))
  occ, expd = torch.from_numpy(g['occ_logits']).cuda(), torch.from_numpy(g['expd_logits']).cuda()
  vis = live.postprocess_occlusions(occ, expd)
  assert vis.dtype == torch.bool
  np.testing.assert_array_equal(vis.cpu().numpy(), g['visible'])
  gen = torch.Generator().manual_seed(9)
  occ = torch.randn(64, 1000, generator=gen) * 4
  expd = torch.randn(64, 1000, generator=gen) * 4
  ref = io_oracle.postprocess_occlusions(occ, expd)
  val = (1 - torch.sigmoid(occ)) * (1 - torch.sigmoid(expd))
  vis = live.postprocess_occlusions(occ.cuda(), expd.cuda()).cpu()
  decided = (val - 0.5).abs() > 1e-6          # exp() may differ in the last ulp at the boundary
  assert torch.equal(vis[decided], ref[decided])
  assert int(decided.sum()) > 0.99 * occ.numel()


@pytest.mark.parametrize('mode', ['first', 'strided'])
@pytest.mark.parametrize('trackwise', [False, True])
def test_tapvid_metrics_golden(mode, trackwise):
  g = np.load(os.path.join(GOLDEN, 'io_tapvid.npz'))
  m = metrics.compute_tapvid_metrics(g['query_points'], g['gt_occluded'], g['gt_tracks'],
                                     g['pred_occluded'], torch.from_numpy(g['pred_tracks']).cuda(),
                                     mode, get_trackwise_metrics=trackwise)
  prefix = f'{mode}_{"track" if trackwise else "video"}_'
  keys = [k[len(prefix):] for k in g.files if k.startswith(prefix)]
  assert sorted(keys) == sorted(m.keys())
  for k in keys:
    np.testing.assert_array_equal(m[k].cpu().numpy(), g[prefix + k], err_msg=k)  # bit-exact


def test_tapvid_counts_large_vs_oracle():
  rng = np.random.default_rng(2)
  B, N, T = 4, 300, 250
  qp = np.stack([rng.integers(0, T, (B, N)).astype(np.float32),
                 rng.uniform(0, 256, (B, N)).astype(np.float32),
                 rng.uniform(0, 256, (B, N)).astype(np.float32)], -1)
  gt = rng.uniform(0, 256, (B, N, T, 2)).astype(np.float32)
  pred = (gt + rng.normal(0, 6, (B, N, T, 2))).astype(np.float32)
  go = rng.uniform(size=(B, N, T)) < 0.4
  po = rng.uniform(size=(B, N, T)) < 0.4
  for mode in ('first', 'strided'):
    ref = io_oracle.tapvid_counts(qp, go, gt, po, pred, mode)
    out = metrics.tapvid_counts(qp, go, gt, po, torch.from_numpy(pred).cuda(), mode)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
  with pytest.raises(ValueError):
    metrics.tapvid_counts(qp, go, gt, po, torch.from_numpy(pred).cuda(), 'last')


def test_tapvid_fused_logits():
  """pred_logits=(occ, expd) == thresholding with postprocess_occlusions first."""
  rng = np.random.default_rng(4)
  B, N, T = 2, 50, 40
  qp = np.zeros((B, N, 3), np.float32)
  gt = rng.uniform(0, 256, (B, N, T, 2)).astype(np.float32)
  pred = torch.from_numpy((gt + rng.normal(0, 3, gt.shape)).astype(np.float32)).cuda()
  go = rng.uniform(size=(B, N, T)) < 0.3
  occ = torch.from_numpy(rng.normal(0, 3, (B, N, T)).astype(np.float32)).cuda()
  expd = torch.from_numpy(rng.normal(0, 3, (B, N, T)).astype(np.float32)).cuda()
  po = ~live.postprocess_occlusions(occ, expd)
  a = metrics.tapvid_counts(qp, go, gt, po, pred, 'first')
  b = metrics.tapvid_counts(qp, go, gt, None, pred, 'first', pred_logits=(occ, expd))
  assert torch.equal(a, b)


def test_online_helpers_match_model_calls():
  """live.online_model_init / online_model_predict on uint8 frames == the explicit float calls."""
  model, _, _ = get_model(causal=True)
  g = torch.Generator().manual_seed(6)
  frames = torch.randint(0, 256, (1, 1, 256, 256, 3), generator=g, dtype=torch.uint8).cuda()
  pts = torch.tensor([[[0., 40., 50.], [0., 200., 120.], [0., 128., 128.]]]).cuda()
  feats = live.online_model_init(model, frames, pts)
  state = [{k: v.cuda() for k, v in d.items()} for d in model.construct_initial_causal_state(3, 1)]
  tracks, vis, new_state = live.online_model_predict(model, frames, feats, state)
  video = live.preprocess_frames(frames)
  grids = model.get_feature_grids(video, False)
  feats2 = model.get_query_features(video, False, pts, grids)
  r = model.estimate_trajectories((256, 256), False, grids, feats2, None, 64,
                                  causal_context=state, get_causal_context=True)
  assert torch.equal(tracks, r['tracks'][-1])
  assert tracks.shape == (1, 3, 1, 2) and vis.shape == (1, 3, 1) and vis.dtype == torch.bool
  assert len(new_state) == len(state)
