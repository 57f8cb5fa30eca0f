"""Micro-benchmark of tapir_gemm shapes/epilogues (CUDA events, L2 flushed between launches)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tapnet_b200 import _lib  # noqa: E402
from tests import gpu_util as U  # noqa: E402

lib = _lib.load()
dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PEAK = 1422.8


def run(name, M, N, K, P=2, bias=True, gelu=False, resid=False, planes_out=0, f32_out=True, conv=None,
        reps=10, flush_l2=False):
  g = torch.Generator().manual_seed(0)
  if conv is None:
    a = torch.randn(P, M, K, generator=g).to(dev).to(torch.bfloat16)
  else:
    f, h, w, c = conv
    a = torch.randn(P, f, h, w, c, generator=g).to(dev).to(torch.bfloat16)
    M, K = f * h * w, 9 * c
  wt = torch.randn(P, N, K, generator=g).to(dev).to(torch.bfloat16)
  b = torch.randn(N, device=dev) if bias else None
  r = torch.randn(M, N, device=dev) if resid else None
  o32 = torch.empty(M, N, device=dev) if f32_out else None
  opl = torch.empty(planes_out, M, N, dtype=torch.bfloat16, device=dev) if planes_out else None
  lin = U.make_linear(wt, b)
  args = dict()
  def call():
    if conv is None:
      st = lib.tapir_gemm(U.ptr(a), K, M * K, ctypes.byref(lin), M, 0, 0, 0, 0, 0, U.ptr(r), N, int(gelu),
                          U.ptr(o32), N, U.ptr(opl), N, M * N, planes_out, None, 0, 0, U.stream())
    else:
      f, h, w, c = conv
      st = lib.tapir_gemm(U.ptr(a), 0, M * c, ctypes.byref(lin), M, 1, f, h, w, c, U.ptr(r), N, int(gelu),
                          U.ptr(o32), N, U.ptr(opl), N, M * N, planes_out, None, 0, 0, U.stream())
    _lib.check(st, name)
  for _ in range(3):
    call()
  torch.cuda.synchronize()
  evs = []
  for _ in range(reps):
    if flush_l2:
      flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record()
    evs.append((e0, e1))
  torch.cuda.synchronize()
  ms = sorted(a_.elapsed_time(b_) for a_, b_ in evs)[len(evs) // 2]
  fl = 2.0 * M * N * K
  npairs = P * (P + 1) // 2
  print(f'{name:34s} M={M:6d} N={N:5d} K={K:5d} P={P} {ms * 1e3:8.1f} us  alg {fl / ms / 1e9:7.1f} TF/s  '
        f'mma {fl * npairs / ms / 1e9:7.1f} TF/s ({fl * npairs / ms / 1e9 / PEAK:5.1%} of {PEAK})', flush=True)


bn = os.environ.get('TAPIR_B200_BLOCK_N', 'auto')
print('BLOCK_N override:', bn)
run('up: gelu+planes', 12288, 2048, 512, gelu=True, planes_out=2, f32_out=False)
run('up: planes only (no gelu)', 12288, 2048, 512, gelu=False, planes_out=2, f32_out=False)
run('up: gelu + f32 out', 12288, 2048, 512, gelu=True)
run('up: f32 out, no bias', 12288, 2048, 512, bias=False)
run('up x4 rows: gelu+planes', 49152, 2048, 512, gelu=True, planes_out=2, f32_out=False)
run('down: bias+resid f32', 12288, 512, 2048, resid=True)
run('down: f32 only', 12288, 512, 2048, bias=False)
run('down x4 rows', 49152, 512, 2048, resid=True)
run('extra conv 256->1024 gelu+planes', 0, 1024, 0, gelu=True, planes_out=2, f32_out=False, conv=(48, 32, 32, 256))
run('extra conv 1024->256 resid', 0, 256, 0, resid=True, conv=(48, 32, 32, 1024))
run('resnet conv 64->64 @128', 0, 64, 0, bias=False, conv=(48, 128, 128, 64))
run('resnet conv 128->128 @64', 0, 128, 0, bias=False, conv=(48, 64, 64, 128))
run('resnet conv 256->256 @32', 0, 256, 0, bias=False, conv=(48, 32, 32, 256))
run('linear_in 576->512', 12288, 512, 576)
run('cost volume c2 (256 q x 48 f)', 256, 49152, 256, P=3, bias=False)
run('cost volume c4/GPU (512 q x 96 f)', 512, 98304, 256, P=3, bias=False)
run('cost volume 4096 q x 96 f', 4096, 98304, 256, P=3, bias=False, reps=5)
run('P=1 up: gelu+planes(1)', 12288, 2048, 512, P=1, gelu=True, planes_out=1, f32_out=False)
run('P=1 big square', 8192, 8192, 8192, P=1, bias=False, reps=5)
run('P=2 big square', 8192, 8192, 4096, P=2, bias=False, reps=5)
