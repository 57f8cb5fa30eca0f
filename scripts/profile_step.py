"""Runs `--warm` warm-up steps + `--steps` steps of the bench workload (for ncu captures)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from tapnet_b200 import tapir_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--warm', type=int, default=2)
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--frames', type=int, default=bench.T_FRAMES)
ap.add_argument('--queries', type=int, default=bench.Q_PER_GPU)
a = ap.parse_args()
bench.T_FRAMES, bench.Q_PER_GPU = a.frames, a.queries
sd, video, queries = bench.build_inputs(1)
model = tapir_model.TAPIR(pyramid_level=1)
model.load_state_dict(sd)
model = model.cuda().eval()
video, queries = video.cuda(), queries.cuda()
for _ in range(a.warm + a.steps):
  out = model(video, queries)
  torch.cuda.synchronize()
print('done', float(out['tracks'].sum()))
