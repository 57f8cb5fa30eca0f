"""End-to-end parity: CUDA path vs the reference's golden outputs and vs the CPU oracle.

Tolerances are BASELINE.json's: tracks 1e-3 px, occlusion / expected_dist logits 1e-4, stage-A
arg-max cell indices bit-exact against the indices the reference itself computed on the same
clip (`stage_a_argmax` in every golden file, recorded by oracle/make_golden.py).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import synth  # noqa: E402
from oracle import tapir_oracle as O  # noqa: E402
from tapnet_b200 import tapir_model  # noqa: E402
from tests import gpu_util as U  # noqa: E402
from tests.conftest import load_golden  # noqa: E402
from tests.test_stages_gpu import get_model  # noqa: E402

TRACK_TOL = 1e-3
LOGIT_TOL = 1e-4


def _argmax_check(name, model, g):
  """Stage-A arg-max cells of the last estimate_trajectories call vs the reference's own."""
  got = model.last_stage_a_argmax[0].cpu().numpy()
  want = g['stage_a_argmax']
  mism = int((got != want).sum())
  U.record(f'argmax_{name}', maps=int(want.size), mismatches=mism,
           min_top2_margin=float(g['stage_a_margin'].min()))
  assert mism == 0, f'{mism}/{want.size} stage-A arg-max cells differ from the reference'


def _inputs(meta):
  video = synth.make_video(meta['T'], meta['H'], meta['W'], seed=meta['video_seed'])
  q = synth.make_queries(meta['N'], meta['T'], meta['H'], meta['W'], seed=meta['query_seed'],
                         frame0_only=(meta['mode'] == 'causal'))
  return video, q


@pytest.mark.parametrize('name', ['c1_bootstapir_256x8_n16', 'tapir_pl0_noextra_256x4_n8',
                                  'bootstapir_320x384x4_n12', 'causal_256x6_n16',
                                  'tapir_pl2_256x3_n6'])
def test_forward_matches_reference_golden(name):
  g = load_golden(name)
  meta = g['meta']
  kw = meta['model_kwargs']
  model, _, _ = get_model(kw.get('pyramid_level', 1), kw.get('extra_convs', True),
                          kw.get('use_casual_conv', False))
  video, q = _inputs(meta)
  model.capture_stage_a_argmax = True
  out = model(video.cuda(), q.cuda())
  torch.cuda.synchronize()
  model.capture_stage_a_argmax = False
  _argmax_check(name, model, g)
  e_t = np.abs(out['tracks'][0].cpu().numpy() - g['tracks']).max()
  e_o = np.abs(out['occlusion'][0].cpu().numpy() - g['occlusion']).max()
  e_e = np.abs(out['expected_dist'][0].cpu().numpy() - g['expected_dist']).max()
  per_iter = [float(np.abs(t[0].cpu().numpy() - g['tracks_iters'][i]).max())
              for i, t in enumerate(out['unrefined_tracks'])]
  U.record(f'e2e_{name}', tracks_err=e_t, occ_err=e_o, expd_err=e_e, per_iter_tracks=str(per_iter))
  assert e_t < TRACK_TOL and e_o < LOGIT_TOL and e_e < LOGIT_TOL


def test_streaming_matches_reference_golden():
  """pytorch_live_demo.py call pattern: per-frame estimate_trajectories with causal state."""
  g = load_golden('causal_256x6_n16')
  meta = g['meta']
  model, _, _ = get_model(causal=True)
  video, q = _inputs(meta)
  video, q = video.cuda(), q.cuda()
  g0 = model.get_feature_grids(video[:, :1], False)
  qf0 = model.get_query_features(video[:, :1], False, q, g0)
  state = model.construct_initial_causal_state(meta['N'], len(qf0.resolutions) - 1)
  state = [{k: v.cuda() for k, v in d.items()} for d in state]
  tr, oc, ex = [], [], []
  for t in range(meta['T']):
    gr = model.get_feature_grids(video[:, t:t + 1], False)
    r = model.estimate_trajectories((meta['H'], meta['W']), False, gr, qf0, None, 64,
                                    causal_context=state, get_causal_context=True)
    state = r['causal_context']
    tr.append(r['tracks'][-1][0].cpu().numpy())
    oc.append(r['occlusion'][-1][0].cpu().numpy())
    ex.append(r['expected_dist'][-1][0].cpu().numpy())
  e_t = np.abs(np.concatenate(tr, 1) - g['online_tracks']).max()
  e_o = np.abs(np.concatenate(oc, 1) - g['online_occlusion']).max()
  e_e = np.abs(np.concatenate(ex, 1) - g['online_expected_dist']).max()
  e_s = np.abs(state[-1]['block_11_causal_2'][0, :, :, ::64].cpu().numpy() - g['online_state_sub']).max()
  U.record('e2e_streaming', tracks_err=e_t, occ_err=e_o, expd_err=e_e, state_err=e_s)
  assert e_t < TRACK_TOL and e_o < LOGIT_TOL and e_e < LOGIT_TOL and e_s < 5e-4


def test_live_demo_shape_480_two_levels():
  """The README's live-demo shape (480x480, 8 points, refinement levels 256 + 480 = 8 iterations):
  offline-causal forward and per-frame streaming against the reference's golden outputs."""
  g = load_golden('causal_480x3_n8')
  meta = g['meta']
  model, _, _ = get_model(causal=True)
  video, q = _inputs(meta)
  video, q = video.cuda(), q.cuda()
  model.capture_stage_a_argmax = True
  out = model(video, q)
  model.capture_stage_a_argmax = False
  _argmax_check('causal_480x3_n8', model, g)
  e_t = np.abs(out['tracks'][0].cpu().numpy() - g['tracks']).max()
  e_o = np.abs(out['occlusion'][0].cpu().numpy() - g['occlusion']).max()
  e_e = np.abs(out['expected_dist'][0].cpu().numpy() - g['expected_dist']).max()
  assert len(out['unrefined_tracks']) == 8
  assert e_t < TRACK_TOL and e_o < LOGIT_TOL and e_e < LOGIT_TOL
  g0 = model.get_feature_grids(video[:, :1], False)
  qf0 = model.get_query_features(video[:, :1], False, q, g0)
  assert len(qf0.resolutions) == 3
  state = [{k: v.cuda() for k, v in d.items()}
           for d in model.construct_initial_causal_state(meta['N'], len(qf0.resolutions) - 1)]
  tr, oc = [], []
  for t in range(meta['T']):
    gr = model.get_feature_grids(video[:, t:t + 1], False)
    r = model.estimate_trajectories((meta['H'], meta['W']), False, gr, qf0, None, 64,
                                    causal_context=state, get_causal_context=True)
    state = r['causal_context']
    tr.append(r['tracks'][-1][0].cpu().numpy())
    oc.append(r['occlusion'][-1][0].cpu().numpy())
  e_t = np.abs(np.concatenate(tr, 1) - g['online_tracks']).max()
  e_o = np.abs(np.concatenate(oc, 1) - g['online_occlusion']).max()
  e_s = np.abs(state[-1]['block_11_causal_2'][0, :, :, ::64].cpu().numpy() - g['online_state_sub']).max()
  U.record('e2e_live_demo_480', tracks_err=e_t, occ_err=e_o, state_err=e_s)
  assert e_t < TRACK_TOL and e_o < LOGIT_TOL and e_s < 5e-4


def test_forward_1024_three_levels_golden():
  """BASELINE config 5's pyramid (256 / 512 / 1024, 12 iterations) against the reference's golden
  outputs (the enabled c5 test compares with the oracle on a reduced clip)."""
  g = load_golden('bootstapir_1024x2_n6')
  meta = g['meta']
  model, _, _ = get_model()
  video, q = _inputs(meta)
  model.capture_stage_a_argmax = True
  out = model(video.cuda(), q.cuda())
  model.capture_stage_a_argmax = False
  _argmax_check('bootstapir_1024x2_n6', model, g)
  assert len(out['unrefined_tracks']) == 12
  e_t = np.abs(out['tracks'][0].cpu().numpy() - g['tracks']).max()
  e_o = np.abs(out['occlusion'][0].cpu().numpy() - g['occlusion']).max()
  e_e = np.abs(out['expected_dist'][0].cpu().numpy() - g['expected_dist']).max()
  U.record('e2e_1024_three_levels', tracks_err=e_t, occ_err=e_o, expd_err=e_e)
  # tracks are in 1024-pixel units here: the 1e-3 px budget is stated at 256^2
  assert e_t < 4 * TRACK_TOL and e_o < LOGIT_TOL and e_e < LOGIT_TOL


def test_initial_resolution_other_than_256_golden():
  """TAPIR(initial_resolution=(192, 320)) (ctor argument, tapir_model.py:86): 24 x 40 cost map,
  generic-size head; against the reference's golden outputs incl. its arg-max cells."""
  name = 'bootstapir_ir192x320x3_n10'
  g = load_golden(name)
  meta = g['meta']
  sd = synth.make_state_dict(0)
  model = tapir_model.TAPIR(pyramid_level=1, initial_resolution=tuple(meta['model_kwargs']['initial_resolution']))
  model.load_state_dict(sd)
  model = model.cuda().eval()
  video, q = _inputs(meta)
  model.capture_stage_a_argmax = True
  out = model(video.cuda(), q.cuda())
  _argmax_check(name, model, g)
  e_t = np.abs(out['tracks'][0].cpu().numpy() - g['tracks']).max()
  e_o = np.abs(out['occlusion'][0].cpu().numpy() - g['occlusion']).max()
  e_e = np.abs(out['expected_dist'][0].cpu().numpy() - g['expected_dist']).max()
  U.record(f'e2e_{name}', tracks_err=e_t, occ_err=e_o, expd_err=e_e)
  assert e_t < TRACK_TOL and e_o < LOGIT_TOL and e_e < LOGIT_TOL


def test_chunking_and_oracle_agreement_larger():
  """T=12, N=96 against the CPU oracle (a size the oracle finishes in seconds)."""
  model, sd, cfg = get_model()
  T, N = 12, 96
  video, q = synth.make_video(T), synth.make_queries(N, T)
  with torch.no_grad():
    ref = O.forward(sd, cfg, video, q)
  out = model(video.cuda(), q.cuda())
  e_t = (out['tracks'].cpu() - ref['tracks']).abs().max().item()
  e_o = (out['occlusion'].cpu() - ref['occlusion']).abs().max().item()
  e_e = (out['expected_dist'].cpu() - ref['expected_dist']).abs().max().item()
  U.record('e2e_oracle_T12_N96', tracks_err=e_t, occ_err=e_o, expd_err=e_e)
  assert e_t < TRACK_TOL and e_o < LOGIT_TOL and e_e < LOGIT_TOL


def test_precision_modes():
  """precision='bf16x6' (three bf16 terms, six MMAs) is fp32-equivalent; 'bf16' runs but is
  outside the parity budget by design."""
  g = load_golden('c1_bootstapir_256x8_n16')
  meta = g['meta']
  video, q = _inputs(meta)
  m6, _, _ = get_model(precision='bf16x6')
  out = m6(video.cuda(), q.cuda())
  e_t = np.abs(out['tracks'][0].cpu().numpy() - g['tracks']).max()
  e_o = np.abs(out['occlusion'][0].cpu().numpy() - g['occlusion']).max()
  e_e = np.abs(out['expected_dist'][0].cpu().numpy() - g['expected_dist']).max()
  m1, _, _ = get_model(precision='bf16')
  out1 = m1(video.cuda(), q.cuda())
  e1 = np.abs(out1['occlusion'][0].cpu().numpy() - g['occlusion']).max()
  U.record('e2e_precision_modes', x6_tracks=e_t, x6_occ=e_o, x6_expd=e_e, x1_occ=e1)
  assert e_t < 5e-4 and e_o < 3e-5 and e_e < 3e-5
  assert torch.isfinite(out1['tracks']).all()
