"""The frame gather fused into the render kernel (swb_step_render_gather + PeerFrames).

The gathered buffer is compared with the CPU oracle's frames of every rank (not with a second
engine).  Two ranks share cuda:0 (the GPU tier of the test-suite has one device): each maps the other's
gathered buffer through CUDA IPC and its render kernel stores every frame into both.  NCCL
refuses two ranks on one device, so the handle exchange and the completion barrier run on
gloo (PeerFrames(host_barrier=True)); the kernel path is the one bench.py uses at N > 1.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

E, STEPS = 96, 5


def _worker(rank, world, port, mode, ret):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from spriteworld_b200 import distributed, workloads
    torch.cuda.set_device(0)
    wl = workloads.WORKLOADS['c4']()
    K = 4
    from tests import fixtures
    eng, raster, scenes = workloads.build_engine(wl, E, K, device=0, seed=1000 + rank)
    bo = fixtures.workload_oracle(wl, scenes, E, K)      # CPU oracle over this rank's scenes
    acts_np = wl.sample_actions(np.random.RandomState(7 + rank), STEPS, E)
    acts = torch.from_numpy(acts_np).cuda()
    peer = distributed.PeerFrames(E, (wl.image_size[1], wl.image_size[0], 3), 'cuda:0',
                                  n_slots=2, host_barrier=True)
    ok = True
    for t in range(STEPS):
      if mode == 'stores':
        res = eng.step_gather(acts[t], raster, peer.slot(t))
        peer.barrier()
      else:   # copy engines: render into this rank's block, push it to the peer
        res = eng.step(acts[t], raster, peer.own_slab(t))
        peer.push(t)
        res.frames = peer.frames[t % 2]
      bo.step(acts_np[t])                                    # the oracle, same scenes and actions
      mine = torch.from_numpy(bo.frames.copy())
      torch.cuda.synchronize()
      ok = ok and np.array_equal(res.step_type.cpu().numpy(), bo.step_type)
      ok = ok and np.allclose(res.reward.cpu().numpy(), bo.reward, rtol=1e-14, atol=1e-13)
      # every rank's ORACLE frames, exchanged on the host, are what the buffer must hold
      parts = [torch.empty((E,) + tuple(mine.shape[1:]), dtype=torch.uint8) for _ in range(world)]
      dist.all_gather(parts, mine)
      ok = ok and torch.equal(res.frames.cpu(), torch.cat(parts))
      dist.barrier()   # nobody overwrites a slot a peer is still comparing
    ret[rank] = bool(ok)
    peer.close()
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('mode', ['stores', 'copy_engine'])
def test_peer_gather_two_ranks_one_device(mode):
  world = 2
  port = 29500 + (os.getpid() + 77 + len(mode)) % 2000
  with mp.Manager() as m:
    ret = m.dict()
    mp.spawn(_worker, args=(world, port, mode, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
