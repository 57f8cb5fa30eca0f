"""Why the tensor-core path is split-bf16: CPU emulation of the MMA arithmetic vs fp32.

(The product never runs this code; it documents the precision decision in DESIGN.md section 2.)
"""
import torch

from oracle import synth
from oracle import tapir_oracle as O


def _errs(planes, sd, cfg, video, q, base):
  out = O.forward(sd, cfg, video, q, ctx=O.Ctx(O.split_bf16_mm(planes)))
  return ((out['tracks'] - base['tracks']).abs().max().item(),
          (out['occlusion'] - base['occlusion']).abs().max().item(),
          (out['expected_dist'] - base['expected_dist']).abs().max().item())


def test_single_pass_bf16_misses_the_budget_and_split_meets_it():
  torch.manual_seed(0)
  sd = synth.make_state_dict(0)
  cfg = O.Config()
  video, q = synth.make_video(3), synth.make_queries(12, 3)
  with torch.no_grad():
    base = O.forward(sd, cfg, video, q)
    t1, o1, e1 = _errs(1, sd, cfg, video, q, base)
    t2, o2, e2 = _errs(2, sd, cfg, video, q, base)
  assert t1 > 1e-3 or o1 > 1e-4 or e1 > 1e-4      # plain bf16 operands: out of budget
  assert t2 < 1e-3 and o2 < 1e-4 and e2 < 1e-4    # hi/lo split, 3 MMAs: inside the budget
