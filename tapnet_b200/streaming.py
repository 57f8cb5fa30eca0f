"""Streaming / online tracking with CUDA-graph replay (BASELINE.json config 3).

`tapnet/pytorch_live_demo.py:44-85` runs, per camera frame,
    feature_grids = model.get_feature_grids(frame)            (one frame)
    model.estimate_trajectories(..., causal_context=state, get_causal_context=True)
which is ~260 small kernel launches: at a few query points the GPU idles between launches
(3.7 ms/frame eager for 8 points).  `OnlineTracker` captures exactly that per-frame step once in
a CUDA graph (fixed frame size and point count) and replays it: same kernels, same results bit for
bit, no per-launch host cost.  The causal state lives in static device buffers that the graph
updates in place; `update_query` mirrors `TAPIR.update_query_features` for re-targeting a point.
"""
from typing import Optional, Tuple

import torch

from tapnet_b200 import live
from tapnet_b200 import tapir_model


class OnlineTracker:
  """Graph-replayed per-frame step of a causal TAPIR model (`use_casual_conv=True`)."""

  def __init__(self, model: 'tapir_model.TAPIR', height: int, width: int, num_points: int,
               uint8_frames: bool = False):
    if not model.use_casual_conv:
      raise ValueError('OnlineTracker needs a causal model: TAPIR(use_casual_conv=True)')
    self.model = model
    self.hw = (int(height), int(width))
    self.num_points = int(num_points)
    self.dev = next(model.parameters()).device
    if self.dev.type != 'cuda':
      raise RuntimeError('OnlineTracker runs on CUDA only')
    # uint8_frames: frames are raw [0, 255] camera frames (what pytorch_live_demo.py:139-141 holds
    # before preprocess_frames); the normalisation runs inside the stem conv's loads
    self._frame_dtype = torch.uint8 if uint8_frames else torch.float32
    self._frame = torch.zeros(1, 1, height, width, 3, dtype=self._frame_dtype, device=self.dev)
    self._graph: Optional[torch.cuda.CUDAGraph] = None
    self._graph_sig = None
    self._pinned = False
    self.query_features = None
    self.state = None
    self._out = None

  # -- pytorch_live_demo.py:44-54 online_model_init
  def init(self, frame: torch.Tensor, query_points: torch.Tensor):
    """frame: [H, W, 3] or [1, 1, H, W, 3] float in [-1, 1]; query_points [N, 3] or [1, N, 3] (t,y,x)."""
    frame = self._as_frame(frame)
    qp = query_points.to(self.dev, torch.float32)
    if qp.dim() == 2:
      qp = qp[None]
    if qp.shape[1] != self.num_points:
      raise ValueError(f'expected {self.num_points} query points, got {qp.shape[1]}')
    grids = self.model.get_feature_grids(frame, False)
    qf = self.model.get_query_features(frame, False, qp, grids)
    # private copies: the graph reads these exact buffers on every replay
    self.query_features = tapir_model.QueryFeatures(
        tuple(t.clone() for t in qf.lowres), tuple(t.clone() for t in qf.hires), qf.resolutions)
    n_res = len(qf.resolutions) - 1
    init = self.model.construct_initial_causal_state(self.num_points, n_res)
    self.state = [{k: v.to(self.dev).clone() for k, v in d.items()} for d in init]
    self._graph = None
    return self.query_features

  def _as_frame(self, frame):
    frame = frame.to(self.dev, self._frame_dtype, non_blocking=True)
    if frame.dim() == 3:
      frame = frame[None, None]
    if tuple(frame.shape[2:4]) != self.hw:
      raise ValueError(f'frame size {tuple(frame.shape[2:4])} != {self.hw}')
    return frame

  def _step_eager(self):
    grids = self.model.get_feature_grids(self._frame, False)
    # single-frame step: the causal state is updated in place (causal_context_out = state)
    r = self.model.estimate_trajectories(self.hw, False, grids, self.query_features, None, 64,
                                         causal_context=self.state, get_causal_context=True,
                                         causal_context_out=self.state)
    tracks = r['tracks'][-1]
    occ, expd = r['occlusion'][-1], r['expected_dist'][-1]
    visibles = live.postprocess_occlusions(occ, expd)  # pytorch_live_demo.py:57-59
    return tracks, visibles, occ, expd

  def _capture(self):
    # warm-up on a side stream (allocates workspaces, packs weights, fills the constant cache);
    # the state is saved and restored so that warm-up steps do not advance it
    saved = [{k: v.clone() for k, v in d.items()} for d in self.state]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
      for _ in range(2):
        self._step_eager()
    torch.cuda.current_stream().wait_stream(s)
    for d, sv in zip(self.state, saved):
      for k in d:
        d[k].copy_(sv[k])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
      self._out = self._step_eager()
    for d, sv in zip(self.state, saved):  # capture does not execute, but be explicit
      for k in d:
        d[k].copy_(sv[k])
    self._graph = g
    # The graph holds RAW pointers into the model's workspaces and packed weight planes.  Pin
    # the workspaces (TAPIR._workspace then parks outgrown buffers instead of freeing them) and
    # remember which buffers / weights the capture saw: step() re-captures when another caller
    # of the same model (a larger offline clip, a second tracker, load_state_dict) replaced them.
    if not self._pinned:
      self.model._ws_pins += 1
      self._pinned = True
    self._graph_sig = self._model_sig()

  def _model_sig(self):
    return (self.model._ws_generation, self.model._param_sig())

  def close(self):
    """Drops the graph and releases the workspace pin."""
    self._graph = None
    if self._pinned:
      self.model._ws_pins -= 1
      self._pinned = False
      if self.model._ws_pins == 0:
        self.model._ws_retired.clear()

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass

  # -- pytorch_live_demo.py:62-85 online_model_predict
  def step(self, frame: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (tracks [1, N, 1, 2] in frame pixels (x, y), visibles [1, N, 1] bool).  The
    returned tensors are the graph's static outputs: copy them if they must outlive the next step."""
    if self.query_features is None:
      raise RuntimeError('call init(frame, query_points) first')
    self._frame.copy_(self._as_frame(frame))
    if self._graph is not None and self._graph_sig != self._model_sig():
      self._graph = None  # stale pointers: the model's buffers or weights changed since capture
    if self._graph is None:
      self._capture()
    self._graph.replay()
    return self._out[0], self._out[1]

  @property
  def last_logits(self):
    """(occlusion, expected_dist) logits of the last step, [1, N, 1] each."""
    return self._out[2], self._out[3]

  def update_query(self, frame: torch.Tensor, query_point: torch.Tensor, idx: int):
    """Re-targets point `idx` at (t, y, x) = query_point on `frame` (pytorch_live_demo.py:188-200):
    its features are re-sampled and its causal state zeroed, in place (graph buffers stay valid)."""
    frame = self._as_frame(frame)
    qp = query_point.to(self.dev, torch.float32).reshape(1, 1, 3)
    grids = self.model.get_feature_grids(frame, False)
    new = self.model.get_query_features(frame, False, qp, grids)
    self.query_features, self.state = self.model.update_query_features(
        self.query_features, new, idx, self.state)
