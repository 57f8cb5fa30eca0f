"""One eager streaming step (causal model, 1024 points) for ncu launch lists / captures."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tapnet_b200 import live, synth, tapir_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--warm', type=int, default=2)
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--points', type=int, default=1024)
a = ap.parse_args()
sd = synth.make_state_dict(0)
m = tapir_model.TAPIR(pyramid_level=1, use_casual_conv=True)
m.load_state_dict(sd)
m = m.cuda().eval()
clip = synth.make_video(4).cuda()
q = synth.make_queries(a.points, 1, frame0_only=True).cuda()
qf = live.online_model_init(m, clip[:, :1], q)
st = [{k: v.cuda() for k, v in d.items()} for d in m.construct_initial_causal_state(a.points, 1)]
for t in range(a.warm + a.steps):
  _, _, st = live.online_model_predict(m, clip[:, t % 4:t % 4 + 1], qf, st)
torch.cuda.synchronize()
print('done')
