"""Pins the CPU oracle (oracle/tapir_oracle.py) to the reference's own outputs.

The fixtures under tests/golden were produced by the unmodified reference
(`python -m oracle.make_golden`).  Tolerances are the reference's own re-chunking /
re-association noise floor (7.6e-5 px, 4.4e-6 logits) with head-room.
"""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle import tapir_oracle as O
from tests.conftest import load_golden

TRACK_TOL = 5e-4
LOGIT_TOL = 5e-5

OFFLINE = ['c1_bootstapir_256x8_n16', 'tapir_pl0_noextra_256x4_n8', 'bootstapir_320x384x4_n12',
           'causal_256x6_n16', 'causal_480x3_n8', 'bootstapir_1024x2_n6',
           'bootstapir_ir192x320x3_n10', 'tapir_pl2_256x3_n6']
CAUSAL = ['causal_256x6_n16', 'causal_480x3_n8']   # the second: live-demo shape, two levels


def _setup(g):
  m = g['meta']
  kw = m['model_kwargs']
  cfg = O.Config(pyramid_level=kw.get('pyramid_level', 1), extra_convs=kw.get('extra_convs', True),
                 use_casual_conv=kw.get('use_casual_conv', False),
                 initial_resolution=tuple(kw.get('initial_resolution', (256, 256))))
  sd = synth.make_state_dict(m['weights_seed'], cfg.pyramid_level, cfg.extra_convs)
  video = synth.make_video(m['T'], m['H'], m['W'], seed=m['video_seed'])
  q = synth.make_queries(m['N'], m['T'], m['H'], m['W'], seed=m['query_seed'],
                         frame0_only=(m['mode'] == 'causal'))
  return cfg, sd, video, q


@pytest.mark.parametrize('name', OFFLINE)
def test_oracle_matches_reference_golden(name):
  g = load_golden(name)
  cfg, sd, video, q = _setup(g)
  with torch.no_grad():
    grids = O.get_feature_grids(sd, cfg, video)
    qf = O.get_query_features(cfg, video.shape, q, grids)
    assert [list(r) for r in grids.resolutions] == g['meta']['resolutions']
    np.testing.assert_allclose(grids.lowres[-1][0, :, ::5, ::7, ::16].numpy(), g['lowres_sub'],
                               atol=2e-5)
    np.testing.assert_allclose(grids.hires[-1][0, :, ::9, ::11, ::16].numpy(), g['hires_sub'],
                               atol=2e-5)
    np.testing.assert_allclose(qf.lowres[-1][0, :, ::8].numpy(), g['qfeat_lowres'], atol=2e-5)
    np.testing.assert_allclose(qf.hires[-1][0, :, ::8].numpy(), g['qfeat_hires'], atol=2e-5)
    tr = O.estimate_trajectories(sd, cfg, video.shape[-3:-1], grids, qf, q, 64)
    # stage-A arg-max cells: the oracle must pick the reference's cell on every heat map
    h0, w0 = grids.lowres[0].shape[2:4]
    scale = torch.tensor([1.0, cfg.initial_resolution[0] / video.shape[2],
                          cfg.initial_resolution[1] / video.shape[3]])
    _, _, _, am, _ = O.tracks_from_cost_volume(sd, cfg, qf.lowres[0], grids.lowres[0], q * scale,
                                               return_debug=True)
    assert np.array_equal(am[0].numpy().astype(np.int32), g['stage_a_argmax']), name
  for i in range(len(tr['tracks'])):
    np.testing.assert_allclose(tr['tracks'][i][0].numpy(), g['tracks_iters'][i], atol=TRACK_TOL)
    np.testing.assert_allclose(tr['occlusion'][i][0].numpy(), g['occlusion_iters'][i],
                               atol=LOGIT_TOL)
    np.testing.assert_allclose(tr['expected_dist'][i][0].numpy(), g['expected_dist_iters'][i],
                               atol=LOGIT_TOL)
  p = cfg.num_pips_iter
  mean_tracks = torch.stack(tr['tracks'][p::p]).mean(0)[0].numpy()
  np.testing.assert_allclose(mean_tracks, g['tracks'], atol=TRACK_TOL)


@pytest.mark.parametrize('name', CAUSAL)
def test_oracle_streaming_matches_reference_golden(name):
  g = load_golden(name)
  cfg, sd, video, q = _setup(g)
  m = g['meta']
  with torch.no_grad():
    g0 = O.get_feature_grids(sd, cfg, video[:, :1])
    qf0 = O.get_query_features(cfg, video[:, :1].shape, q, g0)
    state = O.initial_causal_state(m['N'], len(qf0.resolutions) - 1)
    tr_l, oc_l, ex_l = [], [], []
    for t in range(m['T']):
      gr = O.get_feature_grids(sd, cfg, video[:, t:t + 1])
      r = O.estimate_trajectories(sd, cfg, (m['H'], m['W']), gr, qf0, None, 64,
                                  causal_context=state, get_causal_context=True)
      state = r['causal_context']
      tr_l.append(r['tracks'][-1][0].numpy())
      oc_l.append(r['occlusion'][-1][0].numpy())
      ex_l.append(r['expected_dist'][-1][0].numpy())
  np.testing.assert_allclose(np.concatenate(tr_l, 1), g['online_tracks'], atol=TRACK_TOL)
  np.testing.assert_allclose(np.concatenate(oc_l, 1), g['online_occlusion'], atol=LOGIT_TOL)
  np.testing.assert_allclose(np.concatenate(ex_l, 1), g['online_expected_dist'], atol=LOGIT_TOL)
  np.testing.assert_allclose(state[-1]['block_11_causal_2'][0, :, :, ::64].numpy(),
                             g['online_state_sub'], atol=LOGIT_TOL)
  # structural invariant: online == offline-causal on the whole clip
  # (offline run must also be without the query-frame override, i.e. query_points=None)
  with torch.no_grad():
    gr = O.get_feature_grids(sd, cfg, video)
    off = O.estimate_trajectories(sd, cfg, (m['H'], m['W']), gr, qf0, None, 64)
  np.testing.assert_allclose(np.concatenate(tr_l, 1), off['tracks'][-1][0].numpy(), atol=2e-3)
  np.testing.assert_allclose(np.concatenate(oc_l, 1), off['occlusion'][-1][0].numpy(), atol=2e-4)


def test_default_resolutions():
  assert O.default_resolutions((256, 256), (256, 256)) == [(256, 256)]
  assert O.default_resolutions((480, 480), (256, 256)) == [(256, 256), (480, 480)]
  assert O.default_resolutions((1024, 1024), (256, 256)) == [(256, 256), (512, 512), (1024, 1024)]
  assert O.default_resolutions((240, 240), (256, 256)) == [(256, 256)]
