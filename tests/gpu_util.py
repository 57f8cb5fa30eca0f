"""Helpers for the -m gpu tests: everything goes through the C ABI (tapnet_b200._lib)."""
import atexit
import ctypes
import json
import os

import torch

from tapnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = {}


def _dump_report():
  if not REPORT:
    return
  out = os.path.join(ROOT, 'gpurun_out')
  try:
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, 'test_report.json')
    old = {}
    if os.path.exists(path):
      with open(path) as fh:
        old = json.load(fh)
    old.update(REPORT)
    with open(path, 'w') as fh:
      json.dump(old, fh, indent=1, sort_keys=True)
  except Exception:  # pylint: disable=broad-except
    pass


atexit.register(_dump_report)


def record(name, **vals):
  REPORT[name] = {k: (float(v) if hasattr(v, '__float__') else v) for k, v in vals.items()}
  print(f'[report] {name}: ' + ', '.join(f'{k}={v}' for k, v in REPORT[name].items()))


def ptr(t):
  return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def split(x, planes, kpad=None):
  """fp32 [rows, cols] (cuda) -> bf16 planes [P, rows, kpad] through tapir_split_planes."""
  lib = _lib.load()
  x = x.contiguous().float()
  rows, cols = x.shape
  kpad = kpad or cols
  out = torch.empty(planes, rows, kpad, dtype=torch.bfloat16, device=x.device)
  _lib.check(lib.tapir_split_planes(ptr(x), cols, ptr(out), kpad, rows * kpad, rows, cols, kpad,
                                    planes, stream()), 'split')
  return out


def planes_matmul_fp64(a_pl, b_pl):
  """sum_{i+j<P} A_i @ B_j^T in float64 on the CPU (exact model of the tensor-core math)."""
  a = a_pl.double().cpu()
  b = b_pl.double().cpu()
  P = a.shape[0]
  acc = None
  for i in range(P):
    for j in range(P - i):
      t = a[i] @ b[j].t()
      acc = t if acc is None else acc + t
  return acc


def make_linear(w_pl, bias):
  P, n, k = w_pl.shape
  return _lib.Linear(w=w_pl.data_ptr(), bias=(bias.data_ptr() if bias is not None else None), N=n,
                     K=k, planes=P, k_logical=k)


def gemm(a_pl, w_pl, bias=None, residual=None, gelu=False, out_f32=True, out_planes=0, impl=0,
         conv=None, stats=None, rows_per_frame=0):
  """Runs tapir_gemm. a_pl: [P, M, K] (plain) or [P, F, H, W, C] (conv=(F,H,W,C))."""
  lib = _lib.load()
  P, n, k = w_pl.shape
  dev = w_pl.device
  if conv is None:
    m = a_pl.shape[1]
    lda, aps = a_pl.shape[2], a_pl.shape[1] * a_pl.shape[2]
    f = h = w = c = 0
  else:
    f, h, w, c = conv
    m = f * h * w
    lda, aps = 0, m * c
  lin = make_linear(w_pl, bias)
  o32 = torch.full((m, n), float('nan'), dtype=torch.float32, device=dev) if out_f32 else None
  opl = torch.zeros(out_planes, m, n, dtype=torch.bfloat16, device=dev) if out_planes else None
  st = lib.tapir_gemm(ptr(a_pl), lda, aps, ctypes.byref(lin), m, 1 if conv else 0, f, h, w, c,
                      ptr(residual), n if residual is not None else 0, int(gelu), ptr(o32), n,
                      ptr(opl), n, m * n, out_planes, ptr(stats), rows_per_frame, impl, stream())
  _lib.check(st, 'tapir_gemm')
  torch.cuda.synchronize()
  return o32, opl
