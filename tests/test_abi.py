"""The C-ABI library loads and exports every symbol include/tapir_b200.h declares (no GPU)."""
import ctypes
import os
import re

from tapnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
  with open(os.path.join(ROOT, 'include', 'tapir_b200.h')) as fh:
    src = fh.read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  names = re.findall(r'\b(tapir_[a-z0-9_]+)\s*\(', src)
  return sorted(set(names))


def test_library_exports_every_declared_symbol():
  lib = _lib.load()
  declared = _declared_functions()
  assert len(declared) >= 17
  for name in declared:
    assert hasattr(lib, name), f'{name} declared in the header but not exported'
  # and the ctypes table binds exactly the declared set
  assert sorted(_lib.SIGNATURES) == declared


def test_version_and_error_string():
  lib = _lib.load()
  assert lib.tapir_abi_version() == 2
  assert isinstance(lib.tapir_last_error(), bytes)


def test_workspace_queries_need_no_gpu():
  lib = _lib.load()
  assert lib.tapir_backbone_workspace_bytes(48, 256, 256, 1, 2) > 1 << 30
  base = 12288 * (512 * 4 * 2 + 512 * 2 * 2 + 2048 * 2 * 2)
  assert base + (40 << 20) <= lib.tapir_mixer_workspace_bytes(12288, 2) <= base + (41 << 20)
  assert lib.tapir_cost_volume_workspace_bytes(256, 48, 32, 32, 256) > 256 * 48 * 1024 * 4


def test_struct_sizes_match_header():
  # computed from the C declarations (LP64): guards the ctypes mirror against drift
  assert ctypes.sizeof(_lib.Linear) == 32
  assert ctypes.sizeof(_lib.ResnetBlock) == 3 * 32 + 4 * 8 + 16
  assert ctypes.sizeof(_lib.ExtraBlock) == 16 + 64
  assert ctypes.sizeof(_lib.BackboneWeights) == 8 + 8 * 144 + 5 * 80 + 8
  assert ctypes.sizeof(_lib.HeadWeights) == 80
  assert ctypes.sizeof(_lib.MixerBlock) == 48 + 64
  assert ctypes.sizeof(_lib.MixerWeights) == 64 + 8 + 12 * 112 + 8
  assert ctypes.sizeof(_lib.MixerIO) == 8 + 8 + 16 + 32 + 8 + 8
  assert ctypes.sizeof(_lib.CorrArgs) == 5 * 24 + 24 + 24 + 48 + 24  # 5 correlation levels (ABI 2)
  assert ctypes.sizeof(_lib.UpdateArgs) == 8 + 40 + 48 + 7 * 8


def test_header_is_plain_c_and_ctypes_mirror_matches_the_compiler(tmp_path):
  """include/tapir_b200.h compiles as C99 (no C++ / torch types in the boundary) and every
  ctypes Structure has the size AND field offsets the C compiler gives the header's struct."""
  import shutil
  import subprocess
  import pytest
  gcc = shutil.which('gcc')
  if gcc is None:
    pytest.skip('no gcc')
  pairs = [('tapir_linear', _lib.Linear), ('tapir_resnet_block', _lib.ResnetBlock),
           ('tapir_extra_block', _lib.ExtraBlock), ('tapir_backbone_weights', _lib.BackboneWeights),
           ('tapir_head_weights', _lib.HeadWeights), ('tapir_mixer_block', _lib.MixerBlock),
           ('tapir_mixer_weights', _lib.MixerWeights), ('tapir_mixer_io', _lib.MixerIO),
           ('tapir_corr_level', _lib.CorrLevel), ('tapir_corr_args', _lib.CorrArgs),
           ('tapir_update_args', _lib.UpdateArgs), ('tapir_tapvid_args', _lib.TapvidArgs)]
  lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tapir_b200.h"', 'int main(void) {']
  for cname, cls in pairs:
    lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
    for fname, _ in cls._fields_:
      lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
    lines.append('  printf("\\n");')
  lines += ['  return 0;', '}']
  src = tmp_path / 'abi_probe.c'
  src.write_text('\n'.join(lines))
  exe = tmp_path / 'abi_probe'
  subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'),
                  str(src), '-o', str(exe)], check=True)
  out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
  assert len(out) == len(pairs)
  for line, (cname, cls) in zip(out, pairs):
    name, size, *offs = line.split()
    assert name == cname
    assert int(size) == ctypes.sizeof(cls), cname
    assert [int(o) for o in offs] == [getattr(cls, f).offset for f, _ in cls._fields_], cname


def test_c_host_example_links_and_fails_loudly_without_a_gpu(tmp_path):
  """examples/c_host.c (plain C99, links the .so directly) builds and runs: version and size
  queries work without a GPU; a compute call returns a status + message instead of crashing."""
  import shutil
  import subprocess
  import pytest
  import torch
  gcc = shutil.which('gcc')
  if gcc is None:
    pytest.skip('no gcc')
  _lib.load()  # the library must have been built
  libdir = os.path.dirname(_lib.LIB_PATH)
  exe = tmp_path / 'c_host'
  subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                  os.path.join(ROOT, 'examples', 'c_host.c'), '-L', libdir, '-ltapir_b200',
                  f'-Wl,-rpath,{libdir}', '-o', str(exe)], check=True)
  out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=120).stdout
  lines = dict(l.split(' ', 1) for l in out.strip().splitlines())
  assert lines['abi'] == '2'
  assert int(lines['backbone_ws']) > 1 << 30 and int(lines['mixer_ws']) > 0
  assert lines['bad_args'].startswith('rc=1 ') and 'bad arguments' in lines['bad_args']
  if not torch.cuda.is_available():
    assert lines['host_pointers'].startswith('rc=3 ') and 'CUDA error' in lines['host_pointers']
