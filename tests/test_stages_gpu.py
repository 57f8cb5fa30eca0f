"""Per-stage parity of the CUDA path (through the C ABI) against the CPU oracle."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import synth  # noqa: E402
from oracle import tapir_oracle as O  # noqa: E402
from tapnet_b200 import _lib, tapir_model  # noqa: E402
from tests import gpu_util as U  # noqa: E402

_MODELS = {}


def get_model(pyramid_level=1, extra_convs=True, causal=False, precision='bf16x3'):
  key = (pyramid_level, extra_convs, causal, precision)
  if key not in _MODELS:
    sd = synth.make_state_dict(0, pyramid_level, extra_convs)
    m = tapir_model.TAPIR(pyramid_level=pyramid_level, extra_convs=extra_convs,
                          use_casual_conv=causal, precision=precision)
    m.load_state_dict(sd)
    _MODELS[key] = (m.cuda().eval(), sd,
                    O.Config(pyramid_level=pyramid_level, extra_convs=extra_convs,
                             use_casual_conv=causal))
  return _MODELS[key]


def maxerr(a, b):
  return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


def test_abi_loads():
  lib = _lib.load()
  assert lib.tapir_abi_version() == 2


@pytest.mark.parametrize('shape', [(2, 256, 256), (1, 64, 96)])
def test_backbone(shape):
  T, H, W = shape
  model, sd, cfg = get_model()
  video = synth.make_video(T, H, W)
  with torch.no_grad():
    cfg2 = cfg._replace(initial_resolution=(H, W))
    ref = O.get_feature_grids(sd, cfg2, video, refinement_resolutions=[(H, W)])
  model.initial_resolution = (H, W)
  try:
    got = model.get_feature_grids(video.cuda(), False, [(H, W)])
  finally:
    model.initial_resolution = (256, 256)
  torch.cuda.synchronize()
  e_lo = maxerr(got.lowres[0], ref.lowres[0])
  e_hi = maxerr(got.hires[0], ref.hires[0])
  U.record(f'backbone_{T}x{H}x{W}', lowres_err=e_lo, hires_err=e_hi)
  assert e_lo < 2e-4 and e_hi < 2e-4


def test_backbone_resize_path():
  model, sd, cfg = get_model()
  video = synth.make_video(1, 320, 384)
  with torch.no_grad():
    ref = O.get_feature_grids(sd, cfg, video)
  got = model.get_feature_grids(video.cuda(), False)
  assert [tuple(r) for r in got.resolutions] == [tuple(r) for r in ref.resolutions]
  for i in range(len(ref.lowres)):
    e = maxerr(got.lowres[i], ref.lowres[i])
    U.record(f'backbone_resize_level{i}', lowres_err=e, hires_err=maxerr(got.hires[i], ref.hires[i]))
    assert e < 3e-4


def test_query_features():
  model, sd, cfg = get_model()
  T, N = 5, 64
  g = torch.Generator().manual_seed(3)
  lo = torch.randn(1, T, 32, 32, 256, generator=g)
  hi = torch.randn(1, T, 64, 64, 128, generator=g)
  q = synth.make_queries(N, T)
  q[0, 0] = torch.tensor([0.0, 0.0, 0.0])
  q[0, 1] = torch.tensor([T - 1.0, 255.9, 255.9])
  q[0, 2] = torch.tensor([2.5, 100.25, 3.75])  # fractional t (allowed by the API)
  grids = O.Grids((lo,), (hi,), ((256, 256),))
  ref = O.get_query_features(cfg, (1, T, 256, 256, 3), q, grids)
  fg = tapir_model.FeatureGrids((lo.cuda(),), (hi.cuda(),), ((256, 256),))
  got = model.get_query_features(torch.empty(1, T, 256, 256, 3), False, q.cuda(), fg)
  e1, e2 = maxerr(got.lowres[0], ref.lowres[0]), maxerr(got.hires[0], ref.hires[0])
  U.record('query_features', lowres_err=e1, hires_err=e2)
  assert e1 < 1e-5 and e2 < 1e-5


def _unit_grid(T, h, w, c, seed):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(1, T, h, w, c, generator=g)
  return x / x.norm(dim=-1, keepdim=True)


def test_cost_volume_tracks():
  model, sd, cfg = get_model()
  lib = _lib.load()
  pk = model._pack()
  T, N = 6, 40
  grid = _unit_grid(T, 32, 32, 256, 5)
  # queries = features sampled from the grid itself (+ noise) so heat maps are peaked
  g = torch.Generator().manual_seed(6)
  idx = torch.randint(0, T * 32 * 32, (N,), generator=g)
  qf = grid.reshape(-1, 256)[idx] + 0.05 * torch.randn(N, 256, generator=g)
  qf = (qf / qf.norm(dim=-1, keepdim=True))[None]
  qp = synth.make_queries(N, T)
  with torch.no_grad():
    pts, occ, expd, am, cv = O.tracks_from_cost_volume(sd, cfg, qf, grid, qp, return_debug=True)
  dev = 'cuda'
  o_pts = torch.empty(N, T, 2, device=dev)
  o_occ = torch.empty(N, T, device=dev)
  o_exp = torch.empty(N, T, device=dev)
  o_am = torch.empty(N, T, dtype=torch.int32, device=dev)
  nbytes = lib.tapir_cost_volume_workspace_bytes(N, T, 32, 32, 256)
  ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
  qf_d, grid_d, qp_d = qf[0].cuda().contiguous(), grid[0].cuda().contiguous(), qp[0].cuda().contiguous()
  _lib.check(lib.tapir_cost_volume_tracks(
      ctypes.byref(pk['head']), U.ptr(qf_d), U.ptr(grid_d), N, T, 32, 32, 256, U.ptr(qp_d), 20.0, 256,
      256, U.ptr(o_pts), U.ptr(o_occ), U.ptr(o_exp), U.ptr(o_am), U.ptr(ws), nbytes, U.stream()), 'cv')
  torch.cuda.synchronize()
  agree = (o_am.cpu().long() == am[0]).float().mean().item()
  e_p, e_o, e_e = maxerr(o_pts, pts[0]), maxerr(o_occ, occ[0]), maxerr(o_exp, expd[0])
  U.record('cost_volume_tracks', argmax_agree=agree, pts_err=e_p, occ_err=e_o, expd_err=e_e)
  assert agree == 1.0
  assert e_p < 1e-3 and e_o < 1e-4 and e_e < 1e-4


# (pyramid_level, gh, gw, use_last).  24 x 72 is a 3:1 panoramic grid: the reference spaces the x
# samples gw/gh = 3 cells apart (x is normalised by h, utils.py:104), wider than the 16-column
# cell box, so the kernel's per-sample path is exercised (levels 0/1) next to the box path.
@pytest.mark.parametrize('cfg_case', [(1, 32, 32, True), (1, 40, 48, True), (0, 32, 32, False),
                                      (1, 24, 72, True)])
def test_local_corr(cfg_case):
  pyr, gh, gw, use_last = cfg_case
  model, sd, cfg = get_model(pyramid_level=pyr, extra_convs=(pyr == 1))
  lib = _lib.load()
  T, N = 3, 37
  hires = _unit_grid(T, 2 * gh, 2 * gw, 128, 7)
  lowres = _unit_grid(T, gh, gw, 256, 8)
  g = torch.Generator().manual_seed(9)
  pos = torch.rand(1, N, T, 2, generator=g) * 256
  pos[0, 0, 0] = torch.tensor([0.0, 0.0])          # corner: zero padding
  pos[0, 1, 0] = torch.tensor([255.99, 255.99])
  pos[0, 2, 0] = torch.tensor([128.0, 64.0])       # exactly-integer sample positions
  pos[0, 3, 0] = torch.tensor([-20.0, 300.0])      # fully outside
  occ = torch.randn(1, N, T, generator=g)
  expd = torch.randn(1, N, T, generator=g)
  qh = torch.randn(1, N, 128, generator=g)
  ql = torch.randn(1, N, 256, generator=g)
  last = torch.randn(1, N, T, 384, generator=g) if use_last else None
  pyramid = [hires, lowres]
  queries = [qh, ql]
  if pyr:
    pyramid.append(torch.nn.functional.avg_pool3d(lowres, (2, 2, 1), (2, 2, 1)))
    queries.append(ql)
  ref = O.local_correlation(cfg, queries, pyramid, pos, last)  # [1,N,T,49L]
  L = 2 + pyr
  kin = 388 + 49 * L
  kpad = (kin + 63) // 64 * 64
  P = 3
  dev = 'cuda'
  grids_d = [p[0].cuda().contiguous() for p in pyramid]
  if pyr:
    pooled = torch.empty_like(grids_d[2])
    _lib.check(lib.tapir_pool_pyramid(U.ptr(grids_d[1]), T, gh, gw, 256, U.ptr(pooled), U.stream()), 'pool')
    torch.cuda.synchronize()
    e_pool = maxerr(pooled, pyramid[2][0])
    U.record(f'pool_{gh}x{gw}', err=e_pool)
    assert e_pool < 1e-6
    grids_d[2] = pooled
  out = torch.empty(P, N * T, kpad, dtype=torch.bfloat16, device=dev)
  ca = _lib.CorrArgs()
  for i, gd in enumerate(grids_d):
    ca.levels[i].grid = gd.data_ptr()
    ca.levels[i].h, ca.levels[i].w, ca.levels[i].C = gd.shape[1], gd.shape[2], gd.shape[3]
  ca.num_levels, ca.num_points, ca.num_frames = L, N, T
  ca.init_h, ca.init_w, ca.planes = 256, 256, P
  pos_d, occ_d, expd_d = pos[0].cuda().contiguous(), occ[0].cuda().contiguous(), expd[0].cuda().contiguous()
  ca.pos, ca.occ, ca.expd = pos_d.data_ptr(), occ_d.data_ptr(), expd_d.data_ptr()
  if use_last:
    last_d = last[0].cuda().contiguous()
    ca.feat_hi, ca.feat_hi_stride_n, ca.feat_hi_stride_t = last_d.data_ptr(), T * 384, 384
    ca.feat_lo, ca.feat_lo_stride_n, ca.feat_lo_stride_t = last_d.data_ptr() + 512, T * 384, 384
    feats_ref = last[0]
  else:
    qh_d, ql_d = qh[0].cuda().contiguous(), ql[0].cuda().contiguous()
    ca.feat_hi, ca.feat_hi_stride_n, ca.feat_hi_stride_t = qh_d.data_ptr(), 128, 0
    ca.feat_lo, ca.feat_lo_stride_n, ca.feat_lo_stride_t = ql_d.data_ptr(), 256, 0
    feats_ref = torch.cat([qh[0], ql[0]], -1)[:, None].expand(-1, T, -1)
  ca.out_planes, ca.out_plane_stride, ca.ld = out.data_ptr(), N * T * kpad, kpad
  _lib.check(lib.tapir_local_corr(ctypes.byref(ca), U.stream()), 'local_corr')
  torch.cuda.synchronize()
  row = out.float().sum(0).cpu().reshape(N, T, kpad)
  e_corr = (row[..., 388:kin] - ref[0]).abs().max().item()
  e_feat = (row[..., 4:388] - feats_ref).abs().max().item()
  e_head = max(row[..., :2].abs().max().item(), (row[..., 2] - occ[0]).abs().max().item(),
               (row[..., 3] - expd[0]).abs().max().item(), row[..., kin:].abs().max().item())
  U.record(f'local_corr_pyr{pyr}_{gh}x{gw}_last{int(use_last)}', corr_err=e_corr, feat_err=e_feat,
           head_err=e_head, corr_scale=ref.abs().max().item())
  assert e_corr < 2e-5 * max(1.0, ref.abs().max().item()) and e_feat < 1e-6 and e_head < 1e-6


@pytest.mark.parametrize('shape', [(24, 40), (48, 16)])
def test_cost_volume_tracks_any_map_size(shape):
  """initial_resolution != (256, 256) (ctor argument, tapir_model.py:86): the cost map is not
  32 x 32 and the generic-size head runs.  Same checks as the 32 x 32 case."""
  gh, gw = shape
  model, sd, cfg = get_model()
  lib = _lib.load()
  pk = model._pack()
  T, N = 3, 20
  grid = _unit_grid(T, gh, gw, 256, 15)
  g = torch.Generator().manual_seed(16)
  idx = torch.randint(0, T * gh * gw, (N,), generator=g)
  qf = grid.reshape(-1, 256)[idx] + 0.05 * torch.randn(N, 256, generator=g)
  qf = (qf / qf.norm(dim=-1, keepdim=True))[None]
  ih, iw = gh * 8, gw * 8
  qp = synth.make_queries(N, T, ih, iw)
  cfg2 = cfg._replace(initial_resolution=(ih, iw))
  with torch.no_grad():
    pts, occ, expd, am, _ = O.tracks_from_cost_volume(sd, cfg2, qf, grid, qp, return_debug=True)
  dev = 'cuda'
  o_pts = torch.empty(N, T, 2, device=dev)
  o_occ = torch.empty(N, T, device=dev)
  o_exp = torch.empty(N, T, device=dev)
  o_am = torch.empty(N, T, dtype=torch.int32, device=dev)
  nbytes = lib.tapir_cost_volume_workspace_bytes(N, T, gh, gw, 256)
  ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
  qf_d, grid_d, qp_d = qf[0].cuda().contiguous(), grid[0].cuda().contiguous(), qp[0].cuda().contiguous()
  _lib.check(lib.tapir_cost_volume_tracks(
      ctypes.byref(pk['head']), U.ptr(qf_d), U.ptr(grid_d), N, T, gh, gw, 256, U.ptr(qp_d), 20.0, ih,
      iw, U.ptr(o_pts), U.ptr(o_occ), U.ptr(o_exp), U.ptr(o_am), U.ptr(ws), nbytes, U.stream()), 'cv')
  torch.cuda.synchronize()
  agree = (o_am.cpu().long() == am[0]).float().mean().item()
  e_p, e_o, e_e = maxerr(o_pts, pts[0]), maxerr(o_occ, occ[0]), maxerr(o_exp, expd[0])
  U.record(f'cost_volume_tracks_{gh}x{gw}', argmax_agree=agree, pts_err=e_p, occ_err=e_o, expd_err=e_e)
  assert agree == 1.0
  assert e_p < 1e-3 and e_o < 1e-4 and e_e < 1e-4


def _run_mixer(model, x, causal, ctx=None, get_ctx=False):
  lib = _lib.load()
  pk = model._pack()
  n, T, cin = x.shape
  kpad = pk['mixer_in']
  P = model._planes
  dev = 'cuda'
  xp = U.split(x.reshape(n * T, cin).cuda(), P, kpad)
  out = torch.empty(n * T, 388, device=dev)
  io = _lib.MixerIO()
  io.x_planes, io.x_plane_stride, io.ldx = xp.data_ptr(), n * T * kpad, kpad
  io.num_points, io.num_frames, io.causal = n, T, int(causal)
  nb = pk['num_blocks']
  keep = []
  if ctx is not None:
    c1 = [ctx[f'block_{i}_causal_1'].cuda().contiguous() for i in range(nb)]
    c2 = [ctx[f'block_{i}_causal_2'].cuda().contiguous() for i in range(nb)]
    keep += c1 + c2
    io.ctx1_in = (ctypes.c_void_p * nb)(*[t.data_ptr() for t in c1])
    io.ctx2_in = (ctypes.c_void_p * nb)(*[t.data_ptr() for t in c2])
  o1 = o2 = None
  if get_ctx:
    o1 = [torch.empty(n, 2, 512, device=dev) for _ in range(nb)]
    o2 = [torch.empty(n, 2, 2048, device=dev) for _ in range(nb)]
    io.ctx1_out = (ctypes.c_void_p * nb)(*[t.data_ptr() for t in o1])
    io.ctx2_out = (ctypes.c_void_p * nb)(*[t.data_ptr() for t in o2])
  io.out, io.ldo = out.data_ptr(), 388
  nbytes = lib.tapir_mixer_workspace_bytes(n * T, P)
  ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
  _lib.check(lib.tapir_mixer_forward(ctypes.byref(pk['mixer']), ctypes.byref(io), U.ptr(ws), nbytes,
                                     U.stream()), 'mixer')
  torch.cuda.synchronize()
  new_ctx = None
  if get_ctx:
    new_ctx = {}
    for i in range(nb):
      new_ctx[f'block_{i}_causal_1'] = o1[i].cpu()
      new_ctx[f'block_{i}_causal_2'] = o2[i].cpu()
  return out.cpu().reshape(n, T, 388), new_ctx


@pytest.mark.parametrize('T', [1, 5, 24, 50])
def test_mixer_noncausal(T):
  model, sd, cfg = get_model()
  g = torch.Generator().manual_seed(T)
  n = 9
  x = torch.randn(n, T, 535, generator=g)
  with torch.no_grad():
    ref, _ = O.mixer(sd, x, O.Ctx(), False)
  got, _ = _run_mixer(model, x, False)
  e = (got - ref).abs().max().item()
  U.record(f'mixer_noncausal_T{T}', err=e, scale=ref.abs().max().item())
  assert e < 2e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize('T', [1, 2, 7, 30])
def test_mixer_causal_with_context(T):
  model, sd, cfg = get_model(causal=True)
  g = torch.Generator().manual_seed(100 + T)
  n = 6
  x = torch.randn(n, T, 535, generator=g)
  ctx = {}
  for i in range(12):
    ctx[f'block_{i}_causal_1'] = torch.randn(n, 2, 512, generator=g)
    ctx[f'block_{i}_causal_2'] = torch.randn(n, 2, 2048, generator=g)
  with torch.no_grad():
    ref, ref_ctx = O.mixer(sd, x, O.Ctx(), True, {k: v.clone() for k, v in ctx.items()}, True)
    ref0, _ = O.mixer(sd, x, O.Ctx(), True, None, False)
  got, got_ctx = _run_mixer(model, x, True, ctx, True)
  got0, _ = _run_mixer(model, x, True, None, False)
  e = (got - ref).abs().max().item()
  e0 = (got0 - ref0).abs().max().item()
  ec = max((got_ctx[k] - ref_ctx[k]).abs().max().item() for k in ref_ctx)
  U.record(f'mixer_causal_T{T}', err=e, err_noctx=e0, ctx_err=ec, scale=ref.abs().max().item())
  s = max(1.0, ref.abs().max().item())
  assert e < 2e-4 * s and e0 < 2e-4 * s and ec < 2e-4 * s


def test_refine_update():
  lib = _lib.load()
  g = torch.Generator().manual_seed(11)
  n, T = 5, 4
  res = torch.randn(n * T, 388, generator=g)
  pos = torch.rand(n, T, 2, generator=g) * 256
  occ, expd = torch.randn(n, T, generator=g), torch.randn(n, T, generator=g)
  feat = torch.randn(n, T, 384, generator=g)
  ua = _lib.UpdateArgs()
  d = lambda t: t.cuda().contiguous()  # noqa: E731
  res_d, pos_d, occ_d, expd_d, feat_d = d(res), d(pos), d(occ), d(expd), d(feat)
  occ_o, expd_o = torch.empty_like(occ_d), torch.empty_like(expd_d)
  feat_o, trk = torch.empty_like(feat_d), torch.empty_like(pos_d)
  ua.res, ua.ld_res, ua.num_points, ua.num_frames = res_d.data_ptr(), 388, n, T
  ua.init_h, ua.init_w, ua.resize_h, ua.resize_w, ua.video_h, ua.video_w = 256, 256, 320, 384, 480, 640
  ua.feat_hi, ua.feat_hi_stride_n, ua.feat_hi_stride_t = feat_d.data_ptr(), T * 384, 384
  ua.feat_lo, ua.feat_lo_stride_n, ua.feat_lo_stride_t = feat_d.data_ptr() + 512, T * 384, 384
  ua.pos, ua.occ_in, ua.expd_in = pos_d.data_ptr(), occ_d.data_ptr(), expd_d.data_ptr()
  ua.occ_out, ua.expd_out, ua.feat_out, ua.tracks_out = (occ_o.data_ptr(), expd_o.data_ptr(),
                                                          feat_o.data_ptr(), trk.data_ptr())
  _lib.check(lib.tapir_refine_update(ctypes.byref(ua), U.stream()), 'update')
  torch.cuda.synchronize()
  r = res.reshape(n, T, 388)
  new_pos = pos + r[..., :2] * torch.tensor([256.0, 256.0]) / torch.tensor([384.0, 320.0])
  assert maxerr(pos_d, new_pos) < 1e-4
  assert maxerr(trk, new_pos * torch.tensor([640.0, 480.0]) / 256.0) < 1e-3
  assert maxerr(occ_o, occ + r[..., 2]) < 1e-6 and maxerr(expd_o, expd + r[..., 3]) < 1e-6
  assert maxerr(feat_o, feat + r[..., 4:]) < 1e-6
