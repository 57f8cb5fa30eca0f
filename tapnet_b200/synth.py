"""Seeded synthetic weights / video / queries for benchmarks, smoke and
tests.  Input generation only: nothing of the reference's arithmetic lives here.

No checkpoint or dataset is reachable offline, so parity is checked on a deterministic
random state dict (same distribution family as torch's default init: U(-1/sqrt(fan_in),
1/sqrt(fan_in)); norm scales 1+0.1*N(0,1), norm biases 0.1*N(0,1) so that affine handling
is actually exercised) and a seeded random video.  The same torch build runs here and on
the GPU box (same image), so these are reproducible there without shipping 218 MB.
"""
import math

import torch

from tapnet_b200 import schema


def _fan_in(shape):
  f = 1
  for d in shape[1:]:
    f *= d
  return f


def make_state_dict(seed: int = 0, pyramid_level: int = 1, extra_convs: bool = True):
  g = torch.Generator().manual_seed(seed)
  sch = schema.state_dict_schema(pyramid_level, extra_convs)
  sd = {}
  for key, shape in sch.items():
    is_norm = ('bn_' in key) or ('layer_norm' in key)
    if is_norm and key.endswith('.weight'):
      t = 1.0 + 0.1 * torch.randn(shape, generator=g)
    elif is_norm:
      t = 0.1 * torch.randn(shape, generator=g)
    else:
      wshape = shape if key.endswith('.weight') else sch[key[:-len('bias')] + 'weight']
      b = 1.0 / math.sqrt(_fan_in(wshape))
      t = (torch.rand(shape, generator=g) * 2 - 1) * b
    sd[key] = t
  return sd


def make_video(num_frames: int, height: int = 256, width: int = 256, seed: int = 1,
               smooth: bool = True):
  """Video in [-1, 1], shape [1, T, H, W, 3].

  `smooth=True` low-pass filters the noise and adds a slow drift between frames so that
  cost volumes have structure (peaked heat maps) instead of pure white noise.
  """
  g = torch.Generator().manual_seed(seed)
  if not smooth:
    return torch.rand(1, num_frames, height, width, 3, generator=g) * 2 - 1
  base = torch.rand(1, 3, height // 4 + 8, width // 4 + 8, generator=g)
  frames = []
  for t in range(num_frames):
    dx, dy = (t * 3) % 29, (t * 2) % 23
    up = torch.nn.functional.interpolate(
        base, scale_factor=4, mode='bilinear', align_corners=False)
    crop = up[:, :, dy:dy + height, dx:dx + width]
    noise = torch.rand(1, 3, height, width, generator=g)
    frames.append((0.8 * crop + 0.2 * noise).permute(0, 2, 3, 1))
  video = torch.stack(frames, dim=1)
  return (video * 2 - 1).contiguous()


def make_queries(num_points: int, num_frames: int, height: int = 256, width: int = 256,
                 seed: int = 2, frame0_only: bool = False):
  """Query points [1, N, 3] as (t, y, x) in raster coordinates."""
  g = torch.Generator().manual_seed(seed)
  if frame0_only:
    t = torch.zeros(1, num_points, 1)
  else:
    t = torch.randint(0, num_frames, (1, num_points, 1), generator=g).float()
  y = torch.rand(1, num_points, 1, generator=g) * height
  x = torch.rand(1, num_points, 1, generator=g) * width
  return torch.cat([t, y, x], dim=-1)
