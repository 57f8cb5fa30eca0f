"""TEST INFRASTRUCTURE - generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
The fixtures pin `oracle/tapir_oracle.py` (and through it the CUDA path) to the reference's
own outputs on seeded synthetic inputs; inputs are regenerated from the seeds recorded in
each file (`oracle/synth.py`), outputs are stored.
"""
import json
import os

import numpy as np
import torch

from oracle import reference_loader, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          'tests', 'golden')

CASES = {
    # name: (model kwargs, T, N, H, W, mode)
    'c1_bootstapir_256x8_n16': (dict(pyramid_level=1), 8, 16, 256, 256, 'offline'),
    'tapir_pl0_noextra_256x4_n8': (dict(pyramid_level=0, extra_convs=False), 4, 8, 256, 256,
                                   'offline'),
    'bootstapir_320x384x4_n12': (dict(pyramid_level=1), 4, 12, 320, 384, 'offline'),
    'causal_256x6_n16': (dict(pyramid_level=1, use_casual_conv=True), 6, 16, 256, 256,
                         'causal'),
    # BASELINE config 5's pyramid: 1024x1024 -> levels 256 / 512 / 1024, 12 refinement iterations
    'bootstapir_1024x2_n6': (dict(pyramid_level=1), 2, 6, 1024, 1024, 'offline'),
    # the README's live-demo shape (17 fps figure): 480x480, 8 points, two levels = 8 iterations
    'causal_480x3_n8': (dict(pyramid_level=1, use_casual_conv=True), 3, 8, 480, 480, 'causal'),
}


def _np(x):
  return x.detach().cpu().numpy()


def run_case(name):
  kwargs, T, N, H, W, mode = CASES[name]
  sd = synth.make_state_dict(0, kwargs.get('pyramid_level', 1), kwargs.get('extra_convs', True))
  model = reference_loader.build(sd, **kwargs)
  video = synth.make_video(T, H, W, seed=1)
  queries = synth.make_queries(N, T, H, W, seed=2, frame0_only=(mode == 'causal'))
  out = {}
  meta = dict(name=name, model_kwargs=kwargs, T=T, N=N, H=H, W=W, mode=mode,
              weights_seed=0, video_seed=1, query_seed=2, torch=torch.__version__)
  with torch.no_grad():
    grids = model.get_feature_grids(video, is_training=False)
    qf = model.get_query_features(video, False, queries, grids)
    # strided sub-samples of the feature grids / query features (full grids are MBs)
    out['lowres_sub'] = _np(grids.lowres[-1][0, :, ::5, ::7, ::16])
    out['hires_sub'] = _np(grids.hires[-1][0, :, ::9, ::11, ::16])
    out['qfeat_lowres'] = _np(qf.lowres[-1][0, :, ::8])
    out['qfeat_hires'] = _np(qf.hires[-1][0, :, ::8])
    meta['resolutions'] = [list(map(int, r)) for r in grids.resolutions]
    torch.manual_seed(123)  # estimate_trajectories shuffles queries with torch.randperm
    tr = model.estimate_trajectories(video.shape[-3:-1], False, grids, qf, queries,
                                     query_chunk_size=64)
    out['tracks_iters'] = np.stack([_np(t[0]) for t in tr['tracks']])
    out['occlusion_iters'] = np.stack([_np(t[0]) for t in tr['occlusion']])
    out['expected_dist_iters'] = np.stack([_np(t[0]) for t in tr['expected_dist']])
    torch.manual_seed(123)
    fw = model(video, queries)
    out['tracks'] = _np(fw['tracks'][0])
    out['occlusion'] = _np(fw['occlusion'][0])
    out['expected_dist'] = _np(fw['expected_dist'][0])
    if mode == 'causal':
      # streaming: frame by frame with causal state (pytorch_live_demo.py:44-85)
      g0 = model.get_feature_grids(video[:, :1], False)
      qf0 = model.get_query_features(video[:, :1], False, queries, g0)
      state = model.construct_initial_causal_state(N, len(qf0.resolutions) - 1)
      tr_l, oc_l, ex_l = [], [], []
      for t in range(T):
        g = model.get_feature_grids(video[:, t:t + 1], False)
        r = model.estimate_trajectories((H, W), False, g, qf0, None, query_chunk_size=64,
                                        causal_context=state, get_causal_context=True)
        state = r['causal_context']
        tr_l.append(_np(r['tracks'][-1][0]))
        oc_l.append(_np(r['occlusion'][-1][0]))
        ex_l.append(_np(r['expected_dist'][-1][0]))
      out['online_tracks'] = np.concatenate(tr_l, axis=1)
      out['online_occlusion'] = np.concatenate(oc_l, axis=1)
      out['online_expected_dist'] = np.concatenate(ex_l, axis=1)
      out['online_state_sub'] = _np(state[-1]['block_11_causal_2'][0, :, :, ::64])
  out['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
  return out


def main():
  os.makedirs(GOLDEN_DIR, exist_ok=True)
  for name in CASES:
    out = run_case(name)
    path = os.path.join(GOLDEN_DIR, name + '.npz')
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
