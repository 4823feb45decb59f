"""world_size-2 gloo test of the multi-GPU host logic (env sharding + frame gather)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spriteworld_b200 import distributed


def test_env_shard_partitions():
  for total, world in ((65536, 8), (10, 4), (7, 2), (3, 4)):
    blocks = [distributed.env_shard(total, r, world) for r in range(world)]
    assert blocks[0][0] == 0
    for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
      assert s0 + c0 == s1
    assert blocks[-1][0] + blocks[-1][1] == total
    assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _worker(rank, world, port, total, ret):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    start, count = distributed.env_shard(total, rank, world)
    g = distributed.StepGatherer(total, (4, 4, 3), 'cpu')
    ids = torch.arange(start, start + count)
    frames = (ids.view(-1, 1, 1, 1) % 251).to(torch.uint8).expand(count, 4, 4, 3).contiguous()
    out = g.gather(frames, ids.to(torch.float64) * 0.5, (ids % 3).to(torch.int8),
                   (ids % 2).to(torch.uint8))
    expect = torch.arange(total)
    ok = (torch.equal(out.frames[:, 0, 0, 0], (expect % 251).to(torch.uint8)) and
          torch.equal(out.reward, expect.to(torch.float64) * 0.5) and
          torch.equal(out.step_type, (expect % 3).to(torch.int8)) and
          torch.equal(out.success, (expect % 2).to(torch.uint8)))
    ret[rank] = bool(ok)
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('total', [64, 37])
def test_gather_world_size_2(total):
  world = 2
  port = 29500 + (os.getpid() + total) % 2000
  with mp.Manager() as m:
    ret = m.dict()
    mp.spawn(_worker, args=(world, port, total, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
