"""torchrun entry: sharded_forward over all ranks == single-GPU forward (same inputs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import synth  # noqa: E402
from tapnet_b200 import distributed as tdist  # noqa: E402
from tapnet_b200 import tapir_model  # noqa: E402


def main():
  rank = int(os.environ['RANK'])
  local = int(os.environ['LOCAL_RANK'])
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  dist.init_process_group('nccl', device_id=dev)
  sd = synth.make_state_dict(0)
  model = tapir_model.TAPIR(pyramid_level=1)
  model.load_state_dict(sd)
  model = model.to(dev).eval()
  T, N = 12, 96
  video, q = synth.make_video(T).to(dev), synth.make_queries(N, T).to(dev)
  out = tdist.sharded_forward(model, video, q, gather_outputs=True)
  ref = model(video, q)
  errs = {k: (out[k] - ref[k]).abs().max().item() for k in ('tracks', 'occlusion', 'expected_dist')}
  ok = errs['tracks'] <= 1e-4 and errs['occlusion'] <= 1e-5 and errs['expected_dist'] <= 1e-5
  # bulk driver: batches dealt round-robin to the ranks == one process doing all of them
  import numpy as np  # noqa: E402
  from tapnet_b200 import bulk  # noqa: E402
  cmodel = tapir_model.TAPIR(pyramid_level=1, use_casual_conv=True)
  cmodel.load_state_dict(sd)
  cmodel = cmodel.to(dev).eval()
  rng = np.random.default_rng(0)
  vids = {'a': rng.integers(0, 256, (5, 64, 80, 3), dtype=np.uint8),
          'b': rng.integers(0, 256, (4, 64, 80, 3), dtype=np.uint8)}
  kw = dict(frame_stride=1, points_per_frame=4, point_batch_size=8, frames_per_step=3)
  shared = bulk.track_many_points(vids, ['a', 'b'], cmodel, **kw)           # 5 batches over the ranks
  alone = bulk.track_many_points(vids, ['a', 'b'], cmodel, group=False, **kw)
  for k in vids:
    e = float(np.abs(shared['separation_tracks'][k] - alone['separation_tracks'][k]).max())
    same_vis = bool((shared['separation_visibility'][k] == alone['separation_visibility'][k]).all())
    errs[f'bulk_{k}'] = e
    ok = ok and e == 0.0 and same_vis and shared['separation_tracks'][k].shape[0] == 36
  t = torch.tensor([1.0 if ok else 0.0], device=dev)
  dist.all_reduce(t, op=dist.ReduceOp.MIN)
  if rank == 0:
    print('errors', errs)
    print('MULTI_GPU_OK' if t.item() == 1.0 else 'MULTI_GPU_FAIL')
  dist.destroy_process_group()
  sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == '__main__':
  main()
