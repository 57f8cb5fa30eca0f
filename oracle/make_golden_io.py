"""TEST INFRASTRUCTURE - generates tests/golden/io_*.npz by executing the reference's own functions.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden_io

`tapnet/pytorch_live_demo.py` opens a camera at import time and `tapnet/tapvid/
evaluation_datasets.py` imports TensorFlow / mediapy (absent here), so the modules cannot be
imported; instead the FunctionDef nodes of the functions on the path are compiled, unmodified,
from the files where they lie and executed with numpy / torch in scope.  Nothing is copied into
this repo: the fixtures hold inputs and the reference's outputs only.
"""
import ast
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import reference_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          'tests', 'golden')


def reference_functions(rel_path, names, scope):
  """Compiles the named top-level functions of a reference file into `scope`."""
  path = os.path.join(reference_loader.REFERENCE_ROOT, rel_path)
  with open(path) as f:
    tree = ast.parse(f.read(), filename=path)
  keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
  assert sorted(n.name for n in keep) == sorted(names), (rel_path, names)
  module = ast.Module(body=keep, type_ignores=[])
  exec(compile(module, path, 'exec'), scope)  # pylint: disable=exec-used
  return [scope[n] for n in names]


class _FakeCapture:
  def __init__(self, image):
    self.image = image

  def read(self):
    return True, self.image


def make_ingest():
  from typing import Mapping, Tuple  # names used by annotations in the reference files
  scope = dict(np=np, torch=torch, F=F, Mapping=Mapping, Tuple=Tuple)
  preprocess_frames, postprocess_occlusions, get_frame = reference_functions(
      'tapnet/pytorch_live_demo.py', ['preprocess_frames', 'postprocess_occlusions', 'get_frame'],
      scope)
  ref_utils = __import__('tapnet.torch.utils', fromlist=['bilinear'])
  rng = np.random.default_rng(11)
  out = {}
  # landscape (odd difference), portrait and square camera frames
  for tag, (h, w), res in (('land', (36, 53), (32, 40)), ('port', (50, 30), (24, 24)),
                           ('square', (24, 24), (24, 24)), ('up', (16, 24), (40, 56))):
    frames = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
    cropped = np.stack([get_frame(_FakeCapture(fr))[1] for fr in frames])
    x = preprocess_frames(torch.from_numpy(cropped))
    y = ref_utils.bilinear(x[None], res)[0]
    out[f'{tag}_frames'] = frames
    out[f'{tag}_cropped_shape'] = np.array(cropped.shape)
    out[f'{tag}_resolution'] = np.array(res)
    out[f'{tag}_preprocessed'] = x.numpy()
    out[f'{tag}_out'] = y.numpy()
  occ = torch.from_numpy(rng.normal(0, 3, size=(5, 40)).astype(np.float32))
  expd = torch.from_numpy(rng.normal(0, 3, size=(5, 40)).astype(np.float32))
  out['occ_logits'] = occ.numpy()
  out['expd_logits'] = expd.numpy()
  out['visible'] = postprocess_occlusions(occ, expd).numpy()
  np.savez_compressed(os.path.join(GOLDEN_DIR, 'io_ingest.npz'), **out)
  print('io_ingest.npz', {k: v.shape for k, v in out.items()})


def make_tapvid():
  from typing import Mapping, Tuple
  scope = dict(np=np, Mapping=Mapping, Tuple=Tuple)
  (compute_tapvid_metrics,) = reference_functions('tapnet/tapvid/evaluation_datasets.py',
                                                  ['compute_tapvid_metrics'], scope)
  rng = np.random.default_rng(5)
  B, N, T = 3, 9, 21
  out = {}
  query_points = np.stack([rng.integers(0, T, size=(B, N)).astype(np.float32) +
                           rng.uniform(-0.4, 0.4, size=(B, N)).astype(np.float32),
                           rng.uniform(0, 256, size=(B, N)).astype(np.float32),
                           rng.uniform(0, 256, size=(B, N)).astype(np.float32)], axis=-1)
  query_points[0, 0, 0] = 2.5   # round-half-to-even cases
  query_points[0, 1, 0] = 3.5
  gt_tracks = rng.uniform(0, 256, size=(B, N, T, 2)).astype(np.float32)
  # errors spread over all five thresholds, some exactly on a threshold
  err = rng.choice([0.3, 0.9, 1.0, 1.7, 3.5, 4.0, 7.0, 12.0, 16.0, 40.0], size=(B, N, T))
  ang = rng.uniform(0, 2 * np.pi, size=(B, N, T))
  pred_tracks = (gt_tracks + np.stack([err * np.cos(ang), err * np.sin(ang)], -1)).astype(np.float32)
  pred_tracks[1, :, ::4] = gt_tracks[1, :, ::4] + np.float32([1.0, 0.0])  # distance exactly 1
  gt_occluded = rng.uniform(size=(B, N, T)) < 0.3
  pred_occluded = np.where(rng.uniform(size=(B, N, T)) < 0.8, gt_occluded, ~gt_occluded)
  gt_occluded[2, 0] = True       # a track never visible -> 0/0 in the trackwise ratios
  out.update(query_points=query_points, gt_occluded=gt_occluded, gt_tracks=gt_tracks,
             pred_occluded=pred_occluded, pred_tracks=pred_tracks)
  for mode in ('first', 'strided'):
    for trackwise in (False, True):
      with np.errstate(divide='ignore', invalid='ignore'):
        m = compute_tapvid_metrics(query_points, gt_occluded, gt_tracks, pred_occluded,
                                   pred_tracks, mode, get_trackwise_metrics=trackwise)
      for k, v in m.items():
        out[f'{mode}_{"track" if trackwise else "video"}_{k}'] = np.asarray(v)
  np.savez_compressed(os.path.join(GOLDEN_DIR, 'io_tapvid.npz'), **out)
  print('io_tapvid.npz', len(out), 'arrays')


if __name__ == '__main__':
  reference_loader.load()  # puts /root/reference and the shims on sys.path
  make_ingest()
  make_tapvid()
