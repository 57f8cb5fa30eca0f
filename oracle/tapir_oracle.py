"""TEST INFRASTRUCTURE - CPU restatement of the TAPIR torch inference path (fp32).

This is the checker, not the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  It restates, as plain
functional torch-CPU code over a `state_dict`, the algorithm of

  * `tapnet/torch/tapir_model.py`  (TAPIR.forward / get_feature_grids / get_query_features /
    estimate_trajectories / refine_pips / tracks_from_cost_volume, lines 139-761)
  * `tapnet/torch/nets.py`          (ResNet/BlockV2 247-427, ExtraConvs 25-89,
    PIPsConvBlock / PIPSMLPMixer 107-244)
  * `tapnet/torch/utils.py`         (bilinear 26-42, map_coordinates_3d 45-73,
    map_coordinates_2d 76-113, soft-argmax 116-193, generate_default_resolutions 275-317)

Pinning: `tests/test_oracle_vs_reference.py` runs the unmodified reference (imported from
/root/reference through `oracle/shims`) on the same seeded inputs and requires agreement;
`oracle/make_golden.py` stores reference outputs under `tests/golden/` so the pin travels to
the GPU box where /root/reference does not exist.

Extra (not in the reference): every dense contraction goes through `ctx.mm`, so tests can
emulate the split-bf16 tensor-core arithmetic of the CUDA path on the CPU.
"""
import math
from typing import Callable, Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


class Config(NamedTuple):
  """Constructor arguments that change the arithmetic (tapir_model.py:73-89)."""
  pyramid_level: int = 1
  extra_convs: bool = True
  use_casual_conv: bool = False
  softmax_temperature: float = 20.0
  num_pips_iter: int = 4
  initial_resolution: Tuple[int, int] = (256, 256)
  feature_extractor_chunk_size: int = 10
  num_mixer_blocks: int = 12


class Grids(NamedTuple):
  lowres: Sequence[torch.Tensor]
  hires: Sequence[torch.Tensor]
  resolutions: Sequence[Tuple[int, int]]


# --------------------------------------------------------------------------------------
# utils.py restatements
# --------------------------------------------------------------------------------------


def default_resolutions(full_size, train_size, num_levels=None):
  """utils.py:275-317 - log-spaced refinement resolutions, multiples of 8."""
  full_size = tuple(int(v) for v in full_size)
  train_size = tuple(int(v) for v in train_size)
  if full_size == train_size:
    return [train_size]
  if num_levels is None:
    ratio = np.array(full_size) / np.array(train_size)
    num_levels = int(np.ceil(np.max(np.log2(ratio))) + 1)
  if num_levels <= 1:
    return [train_size]
  h, w = full_size
  lh, lw = train_size
  out = []
  for i in range(num_levels):
    e = i / (num_levels - 1)
    out.append((int(round((lh * (h / lh) ** e) // 8)) * 8,
                int(round((lw * (w / lw) ** e) // 8)) * 8))
  return out


def resize_video(video, resolution):
  """utils.py:26-42 - bilinear, align_corners=False, over [B,T,H,W,C]."""
  b, t, h, w, c = video.shape
  x = video.permute(0, 1, 4, 2, 3).reshape(b, t * c, h, w)
  x = F.interpolate(x, size=tuple(resolution), mode='bilinear', align_corners=False)
  return x.reshape(b, t, c, *x.shape[-2:]).permute(0, 1, 3, 4, 2)


def scale_coords(coords, in_size, out_size):
  """utils.py:228-231 - `coords * out / in`, evaluated in that order."""
  o = torch.tensor(tuple(float(v) for v in out_size), device=coords.device)
  i = torch.tensor(tuple(float(v) for v in in_size), device=coords.device)
  return coords * o / i


def sample_3d(feats, tyx):
  """utils.py:45-73 - trilinear sample of [B,T,H,W,C] at (t,y,x) grid coordinates.

  t gets +0.5 (so integer t hits the frame exactly), each axis is divided by its own
  size, border padding.
  """
  x = feats.permute(0, 4, 1, 2, 3)
  y = tyx[:, :, None, None, :].float().clone()
  y[..., 0] += 0.5
  y = 2 * (y / torch.tensor(x.shape[2:], dtype=torch.float32)) - 1
  y = torch.flip(y, dims=(-1,))
  out = F.grid_sample(x, y, mode='bilinear', align_corners=False, padding_mode='border')
  return out.squeeze(dim=(3, 4)).permute(0, 2, 1)


def sample_2d(feats, yx):
  """utils.py:76-113 - bilinear gather, zeros padding; BOTH axes normalised by h (quirk)."""
  n, t, h, w, c = feats.shape
  x = feats.permute(0, 1, 4, 2, 3).reshape(n * t, c, h, w)
  n, p, t, s, _ = yx.shape
  y = yx.permute(0, 2, 1, 3, 4).reshape(n * t, p, s, 2)
  y = 2 * (y / h) - 1
  y = torch.flip(y, dims=(-1,)).float()
  out = F.grid_sample(x, y, mode='bilinear', align_corners=False, padding_mode='zeros')
  return out.permute(0, 2, 3, 1).reshape(n, t, p, s, c).permute(0, 2, 1, 3, 4)


def soft_argmax(prob, threshold=5.0):
  """utils.py:116-150 - returns ((x,y) in cell units, flat argmax index)."""
  b, n, t, h, w = prob.shape
  ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
  coords = torch.stack([xs + 0.5, ys + 0.5], dim=-1).float()  # [h,w,2]
  flat = prob.reshape(b, n, t, h * w)
  idx = torch.argmax(flat, dim=-1)
  centre = coords.reshape(-1, 2)[idx]  # [b,n,t,2]
  d2 = ((coords[None, None, None] - centre[:, :, :, None, None, :]) ** 2).sum(-1)
  valid = (d2 < threshold ** 2).float()
  wgt = valid * prob
  num = (coords[None, None, None] * wgt[..., None]).sum(dim=(3, 4))
  den = torch.clamp(wgt.sum(dim=(3, 4)), min=1e-12)[..., None]
  return num / den, idx


# --------------------------------------------------------------------------------------
# arithmetic context (precision emulation hook)
# --------------------------------------------------------------------------------------


class Ctx:
  """`mm(x[M,K], w[N,K]) -> [M,N]`; default = fp32 torch matmul (the reference)."""

  def __init__(self, mm: Optional[Callable] = None):
    self.mm = mm if mm is not None else (lambda x, w: x @ w.t())
    self.use_ctx_conv = mm is not None

  def linear(self, x, w, b=None):
    shp = x.shape
    y = self.mm(x.reshape(-1, shp[-1]), w)
    if b is not None:
      y = y + b
    return y.reshape(*shp[:-1], w.shape[0])

  def conv2d(self, x, w, b=None, stride=1, padding=0):
    if not self.use_ctx_conv:
      return F.conv2d(x, w, b, stride=stride, padding=padding)
    n, c, h, wd = x.shape
    co, ci, kh, kw = w.shape
    cols = F.unfold(x, (kh, kw), padding=padding, stride=stride)  # [n, ci*kh*kw, L]
    ho = (h + 2 * padding - kh) // stride + 1
    wo = (wd + 2 * padding - kw) // stride + 1
    y = self.mm(cols.transpose(1, 2).reshape(-1, ci * kh * kw), w.reshape(co, -1))
    if b is not None:
      y = y + b
    return y.reshape(n, ho, wo, co).permute(0, 3, 1, 2)


def split_bf16_mm(planes: int):
  """Emulates the CUDA GEMM: operands split into `planes` bf16 terms, exact products,
  fp32 accumulation, the (i+j < planes) cross terms only (1, 3 or 6 MMAs)."""

  def split(x):
    parts, r = [], x
    for _ in range(planes):
      p = r.to(torch.bfloat16).float()
      parts.append(p)
      r = r - p
    return parts

  def mm(x, w):
    xs, ws = split(x), split(w)
    acc = None
    for i in range(planes):
      for j in range(planes - i):
        t = xs[i] @ ws[j].t()
        acc = t if acc is None else acc + t
    return acc

  return mm


# --------------------------------------------------------------------------------------
# nets.py restatements
# --------------------------------------------------------------------------------------


def _inorm_relu(x, w, b):
  """nets.py:280-286,315-316 - InstanceNorm2d(affine, eps 1e-5) then ReLU."""
  return torch.relu(F.instance_norm(x, weight=w, bias=b, eps=1e-5))


def resnet(sd, x, ctx):
  """nets.py:417-427 + BlockV2 312-327.  x: [n,3,H,W] -> (unit_1 [n,128,H/4,W/4],
  unit_3 [n,256,H/8,W/8]).  SAME-style asymmetric padding for stride 2."""
  p = 'resnet_torch.'
  x = ctx.conv2d(F.pad(x, (2, 4, 2, 4)), sd[p + 'initial_conv.weight'], stride=2)
  outs = {}
  for g, stride in enumerate((1, 2, 2, 1)):
    for b in range(2):
      q = f'{p}block_groups.{g}.blocks.{b}.'
      s = stride if b == 0 else 1
      y = _inorm_relu(x, sd[q + 'bn_0.weight'], sd[q + 'bn_0.bias'])
      shortcut = x
      if b == 0:
        shortcut = ctx.conv2d(y, sd[q + 'proj_conv.weight'], stride=s)
      pad = (1, 1, 1, 1) if s == 1 else (0, 2, 0, 2)
      y = ctx.conv2d(F.pad(y, pad), sd[q + 'conv_0.weight'], stride=s)
      y = _inorm_relu(y, sd[q + 'bn_1.weight'], sd[q + 'bn_1.bias'])
      y = ctx.conv2d(y, sd[q + 'conv_1.weight'], padding=1)
      x = y + shortcut
    outs[g] = x
  return outs[1], outs[3]


def extra_convs(sd, x, ctx):
  """nets.py:55-62,85-89.  x channel-last [n,h,w,256].  NB the residual is added onto the
  layer-normed x, not onto the block input."""
  for i in range(5):
    q = f'extra_convs.blocks.{i}.'
    x = F.layer_norm(x, (x.shape[-1],), sd[q + 'layer_norm.weight'], sd[q + 'layer_norm.bias'])
    xc = x.permute(0, 3, 1, 2)
    r = ctx.conv2d(xc, sd[q + 'conv.weight'], sd[q + 'conv.bias'], padding=1)
    r = F.gelu(r, approximate='tanh')
    xc = xc + ctx.conv2d(r, sd[q + 'conv_1.weight'], sd[q + 'conv_1.bias'], padding=1)
    x = xc.permute(0, 2, 3, 1)
  return x


def _ln_noshift(x, w):
  return F.layer_norm(x, (x.shape[-1],), w, None)


def _dwconv(x_btc, w, b, causal):
  """Depthwise Conv1d over time; x [B,T,C] -> [B,T',C*mult] (nets.py:121-137,155-172)."""
  x = x_btc.permute(0, 2, 1)
  if causal:
    x = F.pad(x, (2, 0))
    y = F.conv1d(x, w, b, padding=0, groups=x.shape[1])
  else:
    y = F.conv1d(x, w, b, padding=1, groups=x.shape[1])
  return y.permute(0, 2, 1)


def mixer(sd, x, ctx, causal, causal_context=None, get_causal_context=False, num_blocks=12):
  """nets.py:234-244 with PIPsConvBlock.forward 143-186.  x [B,T,Cin] -> [B,T,388]."""
  p = 'torch_pips_mixer.'
  x = ctx.linear(x, sd[p + 'linear.weight'], sd[p + 'linear.bias'])
  new_ctx = {}
  for i in range(num_blocks):
    q = f'{p}blocks.{i}.'
    skip = x
    y = _ln_noshift(x, sd[q + 'layer_norm.weight'])
    n_extra = 0
    if causal_context is not None:
      c1 = causal_context[f'block_{i}_causal_1']
      y = torch.cat([c1, y], dim=-2)
      n_extra = c1.shape[-2]
      new_ctx[f'block_{i}_causal_1'] = y[..., -2:, :]
    h = _dwconv(y, sd[q + 'mlp1_up.weight'], sd[q + 'mlp1_up.bias'], causal)
    h = F.gelu(h, approximate='tanh')
    if causal_context is not None:
      c2 = causal_context[f'block_{i}_causal_2']
      n_extra = c2.shape[-2]
      h = torch.cat([c2, h[..., n_extra:, :]], dim=-2)
      new_ctx[f'block_{i}_causal_2'] = h[..., -2:, :]
    h = _dwconv(h, sd[q + 'mlp1_up_1.weight'], sd[q + 'mlp1_up_1.bias'], causal)
    if causal_context is not None:
      h = h[..., n_extra:, :]
    h = h[..., 0::4] + h[..., 1::4] + h[..., 2::4] + h[..., 3::4]
    x = h + skip
    skip = x
    y = _ln_noshift(x, sd[q + 'layer_norm_1.weight'])
    y = ctx.linear(y, sd[q + 'conv_channels_mixer.mlp2_up.weight'],
                   sd[q + 'conv_channels_mixer.mlp2_up.bias'])
    y = F.gelu(y, approximate='tanh')
    y = ctx.linear(y, sd[q + 'conv_channels_mixer.mlp2_down.weight'],
                   sd[q + 'conv_channels_mixer.mlp2_down.bias'])
    x = y + skip
  x = _ln_noshift(x, sd[p + 'layer_norm.weight'])
  x = ctx.linear(x, sd[p + 'linear_1.weight'], sd[p + 'linear_1.bias'])
  return x, (new_ctx if get_causal_context else {})


# --------------------------------------------------------------------------------------
# tapir_model.py restatements
# --------------------------------------------------------------------------------------


def _l2norm(x):
  """tapir_model.py:370-381."""
  return x / torch.sqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=1e-12))


def _same(r1, r2):
  return all(int(a) == int(b) for a, b in zip(r1, r2))


def get_feature_grids(sd, cfg: Config, video, refinement_resolutions=None, ctx=None) -> Grids:
  """tapir_model.py:293-392."""
  ctx = ctx or Ctx()
  if refinement_resolutions is None:
    refinement_resolutions = default_resolutions(video.shape[2:4], cfg.initial_resolution)
  wanted = [tuple(cfg.initial_resolution)] + [tuple(r) for r in refinement_resolutions]
  lowres, hires, sizes = [], [], []
  curr = (-1, -1)
  latent = hi = None
  shape_hw = None
  for res in wanted:
    if res[0] % 8 != 0 or res[1] % 8 != 0:
      raise ValueError('Image resolution must be a multiple of 8.')
    if not _same(curr, res):
      # quirk (tapir_model.py:337): compares the PREVIOUS resolution with the video size
      if _same(curr, video.shape[-3:-1]):
        vr = video
      else:
        vr = resize_video(video, res)
      curr = res
      n, f, h, w, c = vr.shape
      frames = vr.reshape(n * f, h, w, c).permute(0, 3, 1, 2)
      shape_hw = (h, w)
      cs = cfg.feature_extractor_chunk_size
      step = cs if cs > 0 else frames.shape[0]
      lat_l, hi_l = [], []
      for s in range(0, frames.shape[0], step):
        u1, u3 = resnet(sd, frames[s:s + step], ctx)
        lat_l.append(u3.permute(0, 2, 3, 1))
        hi_l.append(u1.permute(0, 2, 3, 1))
      latent = torch.cat(lat_l, 0)
      hi = torch.cat(hi_l, 0)
      if cfg.extra_convs:
        latent = extra_convs(sd, latent, ctx)
      latent = _l2norm(latent).reshape(n, f, *latent.shape[1:])
      hi = _l2norm(hi).reshape(n, f, *hi.shape[1:])
    lowres.append(latent)
    hires.append(hi)
    sizes.append(tuple(shape_hw))
  return Grids(tuple(lowres), tuple(hires), tuple(sizes))


def get_query_features(cfg: Config, video_shape, query_points, grids: Grids) -> Grids:
  """tapir_model.py:217-291.  video_shape = [B,T,H,W,C]."""
  lo, hi = [], []
  curr = (-1, -1)
  for i, res in enumerate(grids.resolutions):
    if _same(curr, res):
      lo.append(lo[-1])
      hi.append(hi[-1])
      continue
    # NB (tapir_model.py:258-265): curr_resolution is never updated in the reference
    pos = scale_coords(query_points, video_shape[1:4], grids.lowres[i].shape[1:4])
    pos_hi = scale_coords(query_points, video_shape[1:4], grids.hires[i].shape[1:4])
    lo.append(sample_3d(grids.lowres[i], pos))
    hi.append(sample_3d(grids.hires[i], pos_hi))
  return Grids(tuple(lo), tuple(hi), tuple(grids.resolutions))


def tracks_from_cost_volume(sd, cfg: Config, qfeat, grid, query_points, ctx=None,
                            return_debug=False):
  """tapir_model.py:687-761 + utils.heatmaps_to_points 153-193.

  qfeat [B,n,256], grid [B,T,h,w,256], query_points [B,n,3] (t,y,x in initial_resolution
  pixels) or None.  Returns points [B,n,T,2] (x,y in initial_resolution pixels), occlusion
  and expected_dist logits [B,n,T].
  """
  ctx = ctx or Ctx()
  m = 'torch_cost_volume_track_mods.'
  b, t, h, w, c = grid.shape
  n = qfeat.shape[1]
  cv = ctx.mm(grid.reshape(b * t * h * w, c), qfeat.reshape(b * n, c))  # [(b t h w), (b n)]
  assert b == 1
  cv = cv.reshape(t, h, w, n).permute(0, 3, 1, 2).reshape(t * n, 1, h, w)  # (t b n)
  occ = torch.relu(F.conv2d(cv, sd[m + 'hid1.weight'], sd[m + 'hid1.bias'], padding=1))
  pos = F.conv2d(occ, sd[m + 'hid2.weight'], sd[m + 'hid2.bias'], padding=1)
  pos = pos.reshape(t, b, n, h, w).permute(1, 2, 0, 3, 4)  # b n t h w
  prob = F.softmax(pos.reshape(b, n, t, h * w) * cfg.softmax_temperature, dim=-1)
  prob = prob.reshape(b, n, t, h, w)
  pts, idx = soft_argmax(prob)
  ih, iw = cfg.initial_resolution
  pts = scale_coords(pts, (w, h), (iw, ih))
  if query_points is not None:
    qf = torch.round(scale_coords(query_points, (t, ih, iw), (t, h, w))[..., 0:1])
    is_q = (qf == torch.arange(t)[None, None, :])[..., None]
    pts = pts * ~is_q + torch.flip(query_points[:, :, None], dims=(-1,))[..., 0:2] * is_q
  o = F.pad(occ, (0, 2, 0, 2))
  o = torch.relu(F.conv2d(o, sd[m + 'hid3.weight'], sd[m + 'hid3.bias'], stride=2))
  o = o.mean(dim=(-1, -2))
  o = torch.relu(F.linear(o, sd[m + 'hid4.weight'], sd[m + 'hid4.bias']))
  o = F.linear(o, sd[m + 'occ_out.weight'], sd[m + 'occ_out.bias'])
  o = o.reshape(t, b, n, 2).permute(1, 2, 0, 3)
  if return_debug:
    return pts, o[..., 0], o[..., 1], idx, cv.reshape(t, n, h, w)
  return pts, o[..., 0], o[..., 1]


def local_correlation(cfg: Config, queries, pyramid, pos, last_iter):
  """tapir_model.py:599-631 - 7x7 bilinear patches around `pos` at each pyramid level,
  dotted with the query (first iteration of a level) or last-iteration features."""
  ih, iw = cfg.initial_resolution
  corrs = []
  off = torch.arange(-3, 4)
  oy, ox = torch.meshgrid(off, off, indexing='ij')
  ctxo = torch.stack([oy, ox], dim=-1).reshape(-1, 2).float()
  for lvl, (q, grid) in enumerate(zip(queries, pyramid)):
    gh, gw = grid.shape[2], grid.shape[3]
    c = torch.flip(scale_coords(pos, (iw, ih), (gw, gh)), dims=(-1,))  # (y,x)
    c = c.unsqueeze(3) + ctxo[None, None, None]
    nb = sample_2d(grid, c)  # [b,n,t,49,C]
    if last_iter is None:
      corrs.append(torch.einsum('bnfsc,bnc->bnfs', nb, q))
    else:
      lq = last_iter[..., :128] if lvl == 0 else last_iter[..., 128:]
      corrs.append(torch.einsum('bnfsc,bnfc->bnfs', nb, lq))
  return torch.cat(corrs, dim=-1)


def refine_pips(sd, cfg: Config, queries, pyramid, pos, occ, expd, last_iter, resize_hw,
                causal_context, get_causal_context, ctx):
  """tapir_model.py:580-685."""
  ih, iw = cfg.initial_resolution
  rh, rw = resize_hw
  corr = local_correlation(cfg, queries, pyramid, pos, last_iter)
  if last_iter is None:
    feats = torch.cat([queries[0], queries[1]], dim=-1).unsqueeze(2)
    feats = feats.expand(-1, -1, corr.shape[2], -1)
  else:
    feats = last_iter
  x = torch.cat([torch.zeros_like(pos), occ[..., None], expd[..., None], feats, corr], dim=-1)
  b, n, t, cdim = x.shape
  cc = None
  if causal_context is not None:
    cc = {k: v.reshape(b * n, *v.shape[2:]) for k, v in causal_context.items()}
  res, new_cc = mixer(sd, x.reshape(b * n, t, cdim).float(), ctx, cfg.use_casual_conv, cc,
                      get_causal_context, cfg.num_mixer_blocks)
  res = res.reshape(b, n, t, -1)
  if get_causal_context:
    new_cc = {k: v.reshape(b, n, *v.shape[1:]) for k, v in new_cc.items()}
  dpos = scale_coords(res[..., :2], (rw, rh), (iw, ih))
  return (dpos + pos, res[..., 2] + occ, res[..., 3] + expd, res[..., 4:] + feats, new_cc)


def estimate_trajectories(sd, cfg: Config, video_size, grids: Grids, qfeats: Grids,
                          query_points, query_chunk_size=64, causal_context=None,
                          get_causal_context=False, ctx=None, perm=None, debug=None):
  """tapir_model.py:394-578.  `perm=None` uses the identity permutation (the reference
  shuffles with torch.randperm when not causal; results are permutation invariant up to fp
  reassociation)."""
  ctx = ctx or Ctx()
  ih, iw = cfg.initial_resolution
  vh, vw = int(video_size[0]), int(video_size[1])
  num_iters = cfg.num_pips_iter * (len(grids.lowres) - 1)
  nq = qfeats.lowres[0].shape[1]
  if perm is None:
    perm = torch.arange(nq)
  inv = torch.zeros_like(perm)
  inv[perm] = torch.arange(nq)
  if query_chunk_size is None:
    query_chunk_size = nq
  occ_it = [[] for _ in range(num_iters + 1)]
  pts_it = [[] for _ in range(num_iters + 1)]
  exp_it = [[] for _ in range(num_iters + 1)]
  cc_it = [[] for _ in range(num_iters)]

  def to_video(p):
    return scale_coords(p, (iw, ih), (vw, vh))

  nf = grids.lowres[0].shape[1]
  for ch in range(0, nq, query_chunk_size):
    sel = perm[ch:ch + query_chunk_size]
    qp = None
    if query_points is not None:
      qp = scale_coords(query_points[:, sel], (nf, vh, vw), (nf, ih, iw))
    cc_chunk = None
    if causal_context is not None:
      cc_chunk = [{k: v[:, sel] for k, v in d.items()} for d in causal_context]
    if debug is not None:
      pts, occ, expd, idx, _ = tracks_from_cost_volume(
          sd, cfg, qfeats.lowres[0][:, sel], grids.lowres[0], qp, ctx, return_debug=True)
      debug.setdefault('argmax', []).append(idx)
    else:
      pts, occ, expd = tracks_from_cost_volume(
          sd, cfg, qfeats.lowres[0][:, sel], grids.lowres[0], qp, ctx)
    pts_it[0].append(to_video(pts))
    occ_it[0].append(occ)
    exp_it[0].append(expd)
    occ0, expd0 = occ, expd
    feats = None
    for i in range(num_iters):
      lvl = i // cfg.num_pips_iter + 1
      queries = [qfeats.hires[lvl][:, sel], qfeats.lowres[lvl][:, sel]]
      pyramid = [grids.hires[lvl], grids.lowres[lvl]]
      for _ in range(cfg.pyramid_level):
        queries.append(queries[-1])
        pyramid.append(F.avg_pool3d(pyramid[-1], kernel_size=(2, 2, 1), stride=(2, 2, 1)))
      cc = cc_chunk[i] if cc_chunk is not None else None
      pts, occ, expd, feats, new_cc = refine_pips(
          sd, cfg, queries, pyramid, pts, occ, expd, feats, grids.resolutions[lvl], cc,
          get_causal_context, ctx)
      pts_it[i + 1].append(to_video(pts))
      occ_it[i + 1].append(occ)
      exp_it[i + 1].append(expd)
      cc_it[i].append(new_cc)
      if (i + 1) % cfg.num_pips_iter == 0:
        feats = None
        occ, expd = occ0, expd0
  out = dict(
      occlusion=[torch.cat(v, 1)[:, inv] for v in occ_it],
      tracks=[torch.cat(v, 1)[:, inv] for v in pts_it],
      expected_dist=[torch.cat(v, 1)[:, inv] for v in exp_it],
  )
  if get_causal_context:
    out['causal_context'] = [
        {k: torch.cat([d[k] for d in lst], 1)[:, inv] for k in lst[0]} for lst in cc_it]
  if debug is not None and 'argmax' in debug:
    debug['argmax'] = torch.cat(debug['argmax'], 1)[:, inv]
  return out


def forward(sd, cfg: Config, video, query_points, query_chunk_size=64,
            refinement_resolutions=None, ctx=None, perm=None, debug=None):
  """tapir_model.py:139-215."""
  ctx = ctx or Ctx()
  grids = get_feature_grids(sd, cfg, video, refinement_resolutions, ctx)
  qf = get_query_features(cfg, video.shape, query_points, grids)
  tr = estimate_trajectories(sd, cfg, video.shape[-3:-1], grids, qf, query_points,
                             query_chunk_size, ctx=ctx, perm=perm, debug=debug)
  p = cfg.num_pips_iter
  return dict(
      occlusion=torch.stack(tr['occlusion'][p::p]).mean(0),
      tracks=torch.stack(tr['tracks'][p::p]).mean(0),
      expected_dist=torch.stack(tr['expected_dist'][p::p]).mean(0),
      unrefined_occlusion=tr['occlusion'][:-1],
      unrefined_tracks=tr['tracks'][:-1],
      unrefined_expected_dist=tr['expected_dist'][:-1],
  )


def initial_causal_state(num_points, num_resolutions=1, num_blocks=12):
  """tapir_model.py:763-772 (same dict object repeated)."""
  d = {}
  for i in range(num_blocks):
    d[f'block_{i}_causal_1'] = torch.zeros(1, num_points, 2, 512)
    d[f'block_{i}_causal_2'] = torch.zeros(1, num_points, 2, 2048)
  return [d] * num_resolutions * 4
