"""Bulk multi-video point tracking, the batch caller of the online path.

Mirrors `track_many_points` and its helpers in `tapnet/robotap/tapir_clustering.py:968-1179`
(same sampling, same argument meaning, same result dictionary):

  1. harvest: on every `frame_stride`-th frame of every video sample `points_per_frame` random
     points inside `sample_box_corners` and extract their query features from that frame alone;
  2. join them into batches of `point_batch_size` points (last batch padded by repeating its last
     frame's points, padding removed from the result);
  3. track every batch through every video with the CAUSAL model, starting each video from a
     zero causal state; keep the final refinement iteration, threshold visibility at 0.5.

What is different is how the work is fed to the GPU.  The reference runs one jitted online step
per (batch, video, frame) - 2048 rows per step - and recomputes the frame's feature grids for every
batch.  Causal convolutions make the online recurrence equal to running the causal model over
`frames_per_step` frames at once with the 2-frame context carried between steps
(nets.py:149-176; tests/test_properties_gpu.py), so here each video's feature grids are computed
once (uint8 frames straight into the stem conv) and every step advances a batch by
`frames_per_step` frames: 24x more rows per launch, no per-frame host work.  With a process group
the batches are dealt round-robin to the ranks (batches are independent; no data-path collective)
and the per-batch results are gathered at the end.
"""
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from tapnet_b200 import live
from tapnet_b200.tapir_model import QueryFeatures, TAPIR, build_model


def query_features_join(feature_list: Sequence[QueryFeatures]) -> QueryFeatures:
  """Concatenates query features along the point axis (tapir_clustering.py:968-978)."""
  lowres = [x.lowres for x in feature_list]
  hires = [x.hires for x in feature_list]
  return QueryFeatures(lowres=tuple(torch.cat(x, dim=1) for x in zip(*lowres)),
                       hires=tuple(torch.cat(x, dim=1) for x in zip(*hires)),
                       resolutions=feature_list[0].resolutions)


def query_features_count(features: QueryFeatures) -> int:
  """Number of points in a query-features structure (tapir_clustering.py:981-983)."""
  return int(features.lowres[0].shape[1])


def predictions_to_tracks_visibility(predictions: Mapping[str, torch.Tensor], single_step=True):
  """tapir_clustering.py:986-1008: tracks [N, (T), 2] and visibility [N, (T)] in [0, 1]."""
  tracks = predictions['tracks'][0]
  occlusion = predictions['occlusion'][0]
  expected_dist = predictions['expected_dist'][0]
  if single_step:
    tracks, occlusion, expected_dist = tracks[:, 0], occlusion[:, 0], expected_dist[:, 0]
  visibility = (1 - torch.sigmoid(occlusion)) * (1 - torch.sigmoid(expected_dist))
  return tracks, visibility


def sample_query_points(videos_shapes: Sequence[Tuple[int, ...]], frame_stride: int,
                        points_per_frame: int, sample_box_corners, seed: int = 42):
  """The reference's sampling (tapir_clustering.py:1046,1063-1080): one
  `uniform(0, 1, [points_per_frame, 3])` draw per selected frame, videos in order, from
  `np.random.seed(42)`.  Returns a list of (video index, frame index, qp [P, 3] (t=0, y, x))."""
  rng = np.random.RandomState(seed)
  x_scl = sample_box_corners[2] - sample_box_corners[0]
  y_scl = sample_box_corners[3] - sample_box_corners[1]
  x_add, y_add = sample_box_corners[0], sample_box_corners[1]
  out = []
  for sv_idx, shape in enumerate(videos_shapes):
    for i in range(0, shape[0], frame_stride):
      qp = (rng.uniform(0.0, 1.0, [points_per_frame, 3])
            * np.array([0.0, shape[1] * y_scl, shape[2] * x_scl])[None, ...]
            + np.array([0.0, shape[1] * y_add, shape[2] * x_add])[None, ...])
      out.append((sv_idx, i, qp))
  return out


def shard_batches(num_batches: int, rank: int, world: int) -> List[int]:
  """Batches are independent: rank r tracks batches r, r + world, ..."""
  return list(range(rank, num_batches, world))


def _as_u8_device(video, dev) -> torch.Tensor:
  t = torch.as_tensor(video)
  if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
    raise ValueError('track_many_points: videos must be uint8 [T, H, W, 3]')
  return t.to(dev).contiguous()


def harvest_query_features(model: TAPIR, video: torch.Tensor, frame_ids: Sequence[int],
                           points: Sequence[np.ndarray], frames_per_call: int = 32) -> QueryFeatures:
  """Query features of `points[k]` ([P, 3], t ignored) taken from frame `frame_ids[k]` of the
  uint8 device video [T, H, W, 3].  Frames are independent in the backbone and an integer t
  samples one frame only, so many frames share one backbone call; the result equals the
  reference's per-frame `online_model_init` calls, points in frame-major order."""
  feats = []
  dev = video.device
  for s in range(0, len(frame_ids), frames_per_call):
    ids = list(frame_ids[s:s + frames_per_call])
    frames = video[torch.as_tensor(ids, device=dev)][None]            # [1, F, H, W, 3] uint8
    qp = np.stack(points[s:s + frames_per_call]).astype(np.float32)    # [F, P, 3]
    qp[..., 0] = np.arange(len(ids), dtype=np.float32)[:, None]        # t = position in this call
    qp_dev = torch.from_numpy(qp.reshape(1, -1, 3)).to(dev)
    grids = model.get_feature_grids(frames, is_training=False)
    feats.append(model.get_query_features(frames, is_training=False, query_points=qp_dev,
                                          feature_grids=grids))
  return query_features_join(feats)


def track_batch_through_video(model: TAPIR, grids_per_step, video_hw, features: QueryFeatures
                              ) -> Tuple[torch.Tensor, torch.Tensor]:
  """Tracks one batch of points through one video given its per-step feature grids; zero causal
  state at the first frame.  Returns (tracks [N, T, 2], visible [N, T] bool)."""
  n = query_features_count(features)
  dev = features.lowres[0].device
  # construct_initial_causal_state (tapir_model.py:763-772) allocated on the device: one dict of
  # zeros shared by every refinement iteration (0.5 GB of pageable host zeros otherwise)
  zeros = {}
  for i in range(model.num_mixer_blocks):
    zeros[f'block_{i}_causal_1'] = torch.zeros(1, n, 2, 512, dtype=torch.float32, device=dev)
    zeros[f'block_{i}_causal_2'] = torch.zeros(1, n, 2, 2048, dtype=torch.float32, device=dev)
  state = [zeros] * (len(features.resolutions) - 1) * 4
  tracks, visible = [], []
  for grids in grids_per_step:
    r = model.estimate_trajectories(video_hw, is_training=False, feature_grids=grids,
                                    query_features=features, query_points_in_video=None,
                                    query_chunk_size=None, causal_context=state,
                                    get_causal_context=True)
    state = r['causal_context']
    tracks.append(r['tracks'][-1][0])
    visible.append(live.postprocess_occlusions(r['occlusion'][-1][0], r['expected_dist'][-1][0]))
  return torch.cat(tracks, dim=1), torch.cat(visible, dim=1)


def track_many_points(separation_videos: Mapping, demo_episode_ids: Sequence, checkpoint_path,
                      frame_stride: int = 4, points_per_frame: int = 8,
                      point_batch_size: int = 2048,
                      sample_box_corners=(0.1, 0.1, 0.9, 0.9), frames_per_step: int = 24,
                      group=None, device: Optional[torch.device] = None) -> Dict:
  """Tracks random points sampled from the videos through all videos.

  Args follow tapir_clustering.py:1023-1044; `checkpoint_path` may also be an already-built
  causal `TAPIR` module.  Extra: `frames_per_step` (frames advanced per launch sequence),
  `group` (torch.distributed group to share the batches over; None = the default group when
  torch.distributed is initialised, False = this process alone), `device`.

  Returns the reference's dictionary, values as numpy arrays: 'separation_visibility'
  {id: [points, T_id] bool}, 'separation_tracks' {id: [points, T_id, 2]}, 'video_shape',
  'query_features', 'demo_episode_ids', 'query_points' [video idx, frame idx, (y, x)].
  """
  if isinstance(checkpoint_path, TAPIR):
    model = checkpoint_path
  else:
    model = build_model(checkpoint_path, use_casual_conv=True)
  if not model.use_casual_conv:
    raise ValueError('Online model requires causal TAPIR training.')  # tapir_clustering.py:895
  if point_batch_size % points_per_frame != 0:
    raise ValueError('point_batch_size must be a multiple of points_per_frame')
  dev = device or next(model.parameters()).device
  if dev.type != 'cuda':
    raise RuntimeError('track_many_points runs on CUDA only (no CPU fallback)')
  rank, world = 0, 1
  if group is False:        # force single-process execution inside an initialised job
    group = None
  elif group is not None or (torch.distributed.is_available()
                             and torch.distributed.is_initialized()):
    rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)

  videos = [_as_u8_device(separation_videos[k], dev) for k in demo_episode_ids]
  shapes = [tuple(int(d) for d in v.shape) for v in videos]
  if len({s[1:] for s in shapes}) != 1:
    raise ValueError('track_many_points: all videos must share one frame size')
  video_hw = shapes[0][1:3]

  # ---- 1. harvest (every rank computes all features: tiny next to the tracking)
  samples = sample_query_points(shapes, frame_stride, points_per_frame, sample_box_corners)
  per_frame: List[QueryFeatures] = []
  for sv_idx, video in enumerate(videos):
    mine = [s for s in samples if s[0] == sv_idx]
    if not mine:
      continue
    joined = harvest_query_features(model, video, [s[1] for s in mine], [s[2] for s in mine])
    for k in range(len(mine)):
      sl = slice(k * points_per_frame, (k + 1) * points_per_frame)
      per_frame.append(QueryFeatures(tuple(t[:, sl] for t in joined.lowres),
                                     tuple(t[:, sl] for t in joined.hires), joined.resolutions))
  out_query_features = query_features_join(per_frame)
  out_query_points = [np.concatenate([np.array([s[0]] * points_per_frame) for s in samples]),
                      np.concatenate([np.array([s[1]] * points_per_frame) for s in samples]),
                      np.concatenate([s[2][..., 1:] for s in samples], axis=0)]

  # ---- 2. batches of point_batch_size points; the last one padded with its last frame
  frames_per_batch = point_batch_size // points_per_frame
  batches, num_extra = [], 0
  for s in range(0, len(per_frame), frames_per_batch):
    chunk = list(per_frame[s:s + frames_per_batch])
    while len(chunk) < frames_per_batch:
      chunk.append(chunk[-1])
      num_extra += points_per_frame
    batches.append(query_features_join(chunk))

  # ---- 3. track: videos outer (feature grids computed once per video), my batches inner
  my_batches = shard_batches(len(batches), rank, world)
  total_frames = sum(s[0] for s in shapes)
  tracks = {b: [] for b in my_batches}
  visible = {b: [] for b in my_batches}
  for video in videos:
    grids_per_step = []
    if my_batches:
      for f0 in range(0, video.shape[0], frames_per_step):
        grids_per_step.append(model.get_feature_grids(video[None, f0:f0 + frames_per_step],
                                                      is_training=False))
    for b in my_batches:
      t, v = track_batch_through_video(model, grids_per_step, video_hw, batches[b])
      tracks[b].append(t)
      visible[b].append(v)
  n_pts = point_batch_size
  all_tracks = torch.zeros(len(batches), n_pts, total_frames, 2, dtype=torch.float32, device=dev)
  all_visible = torch.zeros(len(batches), n_pts, total_frames, dtype=torch.uint8, device=dev)
  for b in my_batches:
    all_tracks[b] = torch.cat(tracks[b], dim=1)
    all_visible[b] = torch.cat(visible[b], dim=1).to(torch.uint8)
  if world > 1:  # every batch was written by exactly one rank, the others hold zeros
    torch.distributed.all_reduce(all_tracks, group=group)
    torch.distributed.all_reduce(all_visible, group=group)
  separation_tracks = all_tracks.reshape(-1, total_frames, 2)
  separation_visibility = all_visible.reshape(-1, total_frames).to(torch.bool)
  pad_start = separation_tracks.shape[0] - num_extra
  separation_tracks = separation_tracks[:pad_start].cpu().numpy()
  separation_visibility = separation_visibility[:pad_start].cpu().numpy()

  bnds, cur = [], 0
  for shp in shapes:
    bnds.append((cur, cur + shp[0]))
    cur += shp[0]
  to_np = lambda t: t.cpu().numpy()
  return {
      'separation_visibility': {k: separation_visibility[:, lb:ub]
                                for k, (lb, ub) in zip(demo_episode_ids, bnds)},
      'separation_tracks': {k: separation_tracks[:, lb:ub]
                            for k, (lb, ub) in zip(demo_episode_ids, bnds)},
      'video_shape': {x: shapes[i] for i, x in enumerate(demo_episode_ids)},
      'query_features': QueryFeatures(tuple(to_np(t) for t in out_query_features.lowres),
                                      tuple(to_np(t) for t in out_query_features.hires),
                                      out_query_features.resolutions),
      'demo_episode_ids': demo_episode_ids,
      'query_points': out_query_points,
  }
