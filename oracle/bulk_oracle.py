"""TEST INFRASTRUCTURE - CPU restatement of the bulk multi-video caller of the online path.

Follows `track_many_points` (tapnet/robotap/tapir_clustering.py:1023-1179) step for step on top
of the restated model (oracle/tapir_oracle.py): per-frame query-feature extraction, batches of
`point_batch_size` points (last one padded), then ONE ONLINE STEP PER FRAME per batch per video
from a zero causal state, final iteration kept, visibility > 0.5.  The reference function itself
needs JAX/Haiku and a checkpoint (parity unpinned for this function as a whole); every model
call it makes is pinned through tapir_oracle's golden files, and the sampling uses numpy's
global-seed stream exactly as the reference does (np.random.seed(42), one uniform draw per frame).
Only tests/ may import this.
"""
import numpy as np
import torch

from oracle import frames_io_oracle as io_oracle
from oracle import tapir_oracle as O


def _join(feature_list):
  return O.Grids(tuple(torch.cat(x, dim=1) for x in zip(*[f.lowres for f in feature_list])),
                 tuple(torch.cat(x, dim=1) for x in zip(*[f.hires for f in feature_list])),
                 feature_list[0].resolutions)


def track_many_points(sd, cfg: O.Config, separation_videos, demo_episode_ids, frame_stride=4,
                      points_per_frame=8, point_batch_size=2048,
                      sample_box_corners=(0.1, 0.1, 0.9, 0.9)):
  assert cfg.use_casual_conv
  np.random.seed(42)                                                       # :1046
  videos = [torch.as_tensor(separation_videos[x]) for x in demo_episode_ids]
  per_frame, samples = [], []
  for sv_idx, sv in enumerate(videos):
    for i in range(0, len(sv), frame_stride):                             # :1063
      x_scl = sample_box_corners[2] - sample_box_corners[0]
      y_scl = sample_box_corners[3] - sample_box_corners[1]
      qp = (np.random.uniform(0.0, 1.0, [points_per_frame, 3])
            * np.array([0.0, sv.shape[1] * y_scl, sv.shape[2] * x_scl])[None]
            + np.array([0.0, sv.shape[1] * sample_box_corners[1],
                        sv.shape[2] * sample_box_corners[0]])[None])      # :1069-1073
      samples.append((sv_idx, i, qp))
      frames = io_oracle.preprocess_frames(sv[None, None, i])             # [1,1,H,W,3]
      grids = O.get_feature_grids(sd, cfg, frames)
      per_frame.append(O.get_query_features(cfg, frames.shape, torch.from_numpy(qp[None]).float(),
                                            grids))
  frames_per_batch = point_batch_size // points_per_frame
  batches, num_extra = [], 0
  for s in range(0, len(per_frame), frames_per_batch):
    chunk = list(per_frame[s:s + frames_per_batch])
    while len(chunk) < frames_per_batch:                                  # :1105-1109
      chunk.append(chunk[-1])
      num_extra += points_per_frame
    batches.append(_join(chunk))
  all_tracks, all_vis = [], []
  for feats in batches:
    n = feats.lowres[0].shape[1]
    tracks, vis = [], []
    for sv in videos:
      state = O.initial_causal_state(n, len(feats.resolutions) - 1)       # zero state per video
      for i in range(len(sv)):
        frames = io_oracle.preprocess_frames(sv[None, None, i])
        grids = O.get_feature_grids(sd, cfg, frames)
        r = O.estimate_trajectories(sd, cfg, frames.shape[-3:-1], grids, feats, None,
                                    query_chunk_size=512, causal_context=state,
                                    get_causal_context=True)
        state = r['causal_context']
        occ, expd = r['occlusion'][-1][0, :, 0], r['expected_dist'][-1][0, :, 0]
        tracks.append(r['tracks'][-1][0, :, 0])
        vis.append(((1 - torch.sigmoid(occ)) * (1 - torch.sigmoid(expd))))
    all_tracks.append(torch.stack(tracks, dim=1))
    all_vis.append(torch.stack(vis, dim=1))
  tracks = torch.cat(all_tracks, dim=0)
  vis = torch.cat(all_vis, dim=0)
  pad_start = tracks.shape[0] - num_extra
  tracks, vis = tracks[:pad_start].numpy(), vis[:pad_start].numpy()
  bnds, cur = [], 0
  for sv in videos:
    bnds.append((cur, cur + sv.shape[0]))
    cur += sv.shape[0]
  joined = _join(per_frame)
  return {
      'separation_visibility_score': {k: vis[:, lb:ub] for k, (lb, ub) in zip(demo_episode_ids, bnds)},
      'separation_visibility': {k: vis[:, lb:ub] > 0.5 for k, (lb, ub) in zip(demo_episode_ids, bnds)},
      'separation_tracks': {k: tracks[:, lb:ub] for k, (lb, ub) in zip(demo_episode_ids, bnds)},
      'video_shape': {x: tuple(videos[i].shape) for i, x in enumerate(demo_episode_ids)},
      'query_features': joined,
      'demo_episode_ids': demo_episode_ids,
      'query_points': [np.concatenate([np.array([s[0]] * points_per_frame) for s in samples]),
                       np.concatenate([np.array([s[1]] * points_per_frame) for s in samples]),
                       np.concatenate([s[2][..., 1:] for s in samples], axis=0)],
  }
