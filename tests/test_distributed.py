"""Host logic of the multi-GPU path with gloo, world_size 2, on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tapnet_b200 import distributed as tdist


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, T, N, ret):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    g = torch.Generator().manual_seed(0)
    full = torch.randn(1, T, 4, 5, 3, generator=g)
    f0, f1, _ = tdist.frame_shard(T, rank, world)
    gathered = tdist.all_gather_frames(full[:, f0:f1].contiguous(), T)
    ok1 = torch.equal(gathered, full)
    outs = torch.randn(1, N, T, 2, generator=g)
    q0, q1 = tdist.query_shard(N, rank, world)
    back = tdist.gather_queries(outs[:, q0:q1].contiguous(), N, 1)
    ok2 = torch.equal(back, outs)
    ret[rank] = bool(ok1 and ok2)
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('T,N', [(6, 8), (5, 7)])
def test_frame_and_query_sharding_roundtrip(T, N):
  world = 2
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(world, _free_port(), T, N, ret), nprocs=world, join=True)
  assert all(ret[r] for r in range(world))


def test_shard_bounds():
  assert tdist.frame_shard(48, 7, 8) == (42, 48, 6)
  assert tdist.frame_shard(5, 1, 2) == (3, 5, 3)
  assert tdist.frame_shard(2, 3, 4)[:2] == (2, 2)  # more ranks than frames: empty slice
  spans = [tdist.query_shard(10, r, 4) for r in range(4)]
  assert spans == [(0, 3), (3, 6), (6, 9), (9, 10)]
