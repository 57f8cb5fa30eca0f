"""The steps either side of the hot path, on the device.

Mirrors the helper functions of `tapnet/pytorch_live_demo.py` (same names and argument meaning),
each one a single kernel of libtapir_b200.so instead of a chain of ATen ops:

  preprocess_frames        :30-41   uint8 [0,255] -> float [-1,1]
  get_frame / crop window  :88-95   centre square crop of a camera frame
  ingest_frames            crop + preprocess + utils.bilinear (torch/utils.py:26-42) in one pass
  postprocess_occlusions   :57-59   visibility flag from the two logits
  online_model_init        :44-54
  online_model_predict     :62-85

`TAPIR.get_feature_grids` also accepts the uint8 frames directly (normalisation fused into the
stem conv), which is what `online_model_*` do here, so a live frame crosses PCIe as uint8 and is
never materialised as float.  CUDA only; raises without the shared library.
"""
import ctypes
from typing import Optional, Tuple

import torch

from tapnet_b200 import _lib


def _ptr(t):
  return ctypes.c_void_p(t.data_ptr())


def _stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, what):
  if t.device.type != 'cuda':
    raise RuntimeError(f'{what}: tensor on {t.device}; tapnet_b200 runs on CUDA only '
                       '(no CPU fallback)')


def center_square_window(height: int, width: int) -> Tuple[int, int, int, int]:
  """(y0, x0, h, w) of the window `get_frame` keeps (pytorch_live_demo.py:88-95): |w-h|//2
  pixels trimmed from both ends of the longer side.  The reference's `[trunc:-trunc]` slice is
  empty when the sides differ by exactly one pixel; that case raises instead."""
  trunc = abs(width - height) // 2
  if width != height and trunc == 0:
    raise ValueError('frame sides differ by one pixel: the reference crop would be empty')
  if width > height:
    return 0, trunc, height, width - 2 * trunc
  if width < height:
    return trunc, 0, height - 2 * trunc, width
  return 0, 0, height, width


def ingest_frames(frames: torch.Tensor, window: Optional[Tuple[int, int, int, int]] = None,
                  resolution: Optional[Tuple[int, int]] = None) -> torch.Tensor:
  """frames: [..., H, W, 3] uint8 on a CUDA device.  Crops to `window` (y0, x0, h, w), normalises
  to [-1, 1] and resizes to `resolution` (h, w) with the reference's bilinear; returns float32
  [..., h, w, 3]."""
  _require_cuda(frames, 'ingest_frames')
  if frames.dtype != torch.uint8 or frames.dim() < 3 or frames.shape[-1] != 3:
    raise ValueError('ingest_frames: expected uint8 frames [..., H, W, 3]')
  frames = frames.contiguous()
  H, W = int(frames.shape[-3]), int(frames.shape[-2])
  lead = tuple(frames.shape[:-3])
  n = 1
  for d in lead:
    n *= int(d)
  y0, x0, ch, cw = (0, 0, H, W) if window is None else (int(v) for v in window)
  oh, ow = (ch, cw) if resolution is None else (int(resolution[0]), int(resolution[1]))
  out = torch.empty(*lead, oh, ow, 3, dtype=torch.float32, device=frames.device)
  if n > 0:
    _lib.check(_lib.load().tapir_ingest_frames(_ptr(frames), n, H, W, y0, x0, ch, cw, _ptr(out),
                                               oh, ow, _stream()), 'tapir_ingest_frames')
  return out


def preprocess_frames(frames: torch.Tensor) -> torch.Tensor:
  """[num_frames, height, width, 3] uint8 -> float32 in [-1, 1] (pytorch_live_demo.py:30-41)."""
  return ingest_frames(frames)


def postprocess_occlusions(occlusions: torch.Tensor, expected_dist: torch.Tensor) -> torch.Tensor:
  """visibles = (1 - sigmoid(occlusions)) * (1 - sigmoid(expected_dist)) > 0.5, bool."""
  _require_cuda(occlusions, 'postprocess_occlusions')
  if occlusions.shape != expected_dist.shape:
    raise ValueError('postprocess_occlusions: shape mismatch')
  occ = occlusions.to(torch.float32).contiguous()
  expd = expected_dist.to(torch.float32).contiguous()
  out = torch.empty(occ.shape, dtype=torch.uint8, device=occ.device)
  if occ.numel() > 0:
    _lib.check(_lib.load().tapir_postprocess_occlusions(_ptr(occ), _ptr(expd), occ.numel(),
                                                        _ptr(out), _stream()),
               'tapir_postprocess_occlusions')
  return out.view(torch.bool)


def online_model_init(model, frames: torch.Tensor, points: torch.Tensor):
  """Query features for `points` [1, N, 3] (t, y, x) on `frames` [1, T, H, W, 3] (uint8 or
  already-normalised float) - pytorch_live_demo.py:44-54."""
  feature_grids = model.get_feature_grids(frames, is_training=False)
  return model.get_query_features(frames, is_training=False, query_points=points,
                                  feature_grids=feature_grids)


def online_model_predict(model, frames: torch.Tensor, features, causal_context):
  """One online step - pytorch_live_demo.py:62-85.  Returns (tracks, visibles, causal_context)
  for the final resolution, like the reference."""
  feature_grids = model.get_feature_grids(frames, is_training=False)
  trajectories = model.estimate_trajectories(
      frames.shape[-3:-1], is_training=False, feature_grids=feature_grids,
      query_features=features, query_points_in_video=None, query_chunk_size=64,
      causal_context=causal_context, get_causal_context=True)
  causal_context = trajectories['causal_context']
  tracks = trajectories['tracks'][-1]
  visibles = postprocess_occlusions(trajectories['occlusion'][-1],
                                    trajectories['expected_dist'][-1])
  return tracks, visibles, causal_context
