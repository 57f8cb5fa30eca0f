"""bench.py's CPU legs: the reference arm must be the UNMODIFIED reference whenever it is importable
(baseline/_ref on the GPU box, /root/reference in the build container), the oracle port otherwise."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tapnet_b200 import synth  # noqa: E402


def test_cpu_arm_runs_the_unmodified_reference_when_importable():
  have_ref = any(os.path.isfile(os.path.join(r, 'tapnet', 'torch', 'tapir_model.py'))
                 for r in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'))
  wl = dict(bench.WORKLOADS['c2'], frames=6)
  video = synth.make_video(6)
  queries = synth.make_queries(32, 6)
  arm = bench.CpuArm(synth.make_state_dict(0), video, queries, wl)
  assert arm.kind == ('reference' if have_ref else 'port')
  if have_ref:
    assert type(arm.model).__module__ == 'tapnet.torch.tapir_model'
  torch.set_num_threads(min(8, os.cpu_count() or 1))
  arm.threads, arm.sweep = torch.get_num_threads(), {}
  sec = arm.sample()
  d = arm.describe(sec)
  assert sec > 0 and d['kind'] == arm.kind and d['unit'] == bench.UNIT
  assert abs(d['value'] - 32 * 6 / sec) < 1e-6 * d['value'] + 0.1
