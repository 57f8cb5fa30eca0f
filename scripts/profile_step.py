"""Runs `--warm` warm-up steps + `--steps` steps of an offline workload (for ncu captures).
Default = the bench headline (BASELINE config 2: 256x256x48, 256 queries)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tapnet_b200 import synth, tapir_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--warm', type=int, default=2)
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--frames', type=int, default=48)
ap.add_argument('--queries', type=int, default=256)
ap.add_argument('--res', type=int, default=256)
a = ap.parse_args()
model = tapir_model.TAPIR(pyramid_level=1)
model.load_state_dict(synth.make_state_dict(0))
model = model.cuda().eval()
video = synth.make_video(a.frames, a.res, a.res, seed=1).cuda()
queries = synth.make_queries(a.queries, a.frames, a.res, a.res, seed=2).cuda()
for _ in range(a.warm + a.steps):
  out = model(video, queries)
  torch.cuda.synchronize()
print('done', float(out['tracks'].sum()))
