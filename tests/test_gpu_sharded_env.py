"""ShardedBatchedEnvironment: one BatchedEnvironment per rank, every rank's step() returns all
envs' outputs (frames stored into every rank's buffer by the render kernel, the 11-byte
per-env records in one all-gather).  Two ranks share cuda:0, so the handle exchange, the
barrier and the record gather run on gloo (host_barrier=True); every rank checks the WHOLE
gathered timestep against the CPU oracles of both shards.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

E_TOTAL, K, STEPS = 128, 4, 60


def _worker(rank, world, port, ret):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from oracle import oracle
    from spriteworld_b200 import constants, environment, scene, sprite_generators
    from spriteworld_b200.configs.cobra import sorting
    from tests import fixtures
    torch.cuda.set_device(0)
    cfg = sorting.get_config('train')
    cfg['max_episode_length'] = 7            # many auto-resets and ring refills in 60 steps
    env = environment.ShardedBatchedEnvironment(E_TOTAL, device=0, seed=5, host_barrier=True,
                                                pool_depth=K, **cfg)
    E = env.n_local
    eng = env.engine
    S = eng.n_slots
    # mirror of this rank's device pool for the oracle, kept in sync by wrapping upload_scenes
    pool = np.zeros((E, K, S), oracle.SPRITE_DTYPE)
    fields = ('x', 'y', 'm00', 'm01', 'm10', 'm11', 'vx', 'vy', 'member', 'shape', 'pos_f32', 'rgb')

    def record(batch, env_ids, ring_slots):
      for f in fields:
        pool[f][np.asarray(env_ids), np.asarray(ring_slots)] = batch[f]

    pushed = eng.upload_scenes

    def upload(batch, env_ids, ring_slots):
      record(batch, env_ids, ring_slots)
      return pushed(batch, env_ids, ring_slots)

    eng.upload_scenes = upload
    rng0 = np.random.RandomState((5 + 7919 * rank) % (2 ** 31))   # the shard's stream
    first = sprite_generators.batch_of(cfg['init_sprites'], E * K, rng0)
    nodes, filters = cfg['task'].compile()
    record(scene.arrays_from_layout(first, S, filters, cfg['renderers']['image'].color_to_rgb),
           np.repeat(np.arange(E), K), np.tile(np.arange(K), E))
    ocfg = fixtures.env_cfg_from_meta(dict(action=cfg['action_space'].compile(), keep_in_frame=True,
                                           max_episode_length=7, nodes=nodes))
    bo = oracle.BatchOracle(ocfg, oracle.shape_table(constants.SHAPES), oracle.raster_cfg(64, 64, 5), pool)
    bo.pool = pool
    rng = np.random.RandomState(11)           # the same global action stream on every rank
    ok, n_first = True, 0
    for t in range(STEPS):
      a = rng.uniform(0, 1, (E_TOTAL, 4)).astype(np.float32)
      ts = env.step(a)                        # global batch: the env takes its shard
      bo.step(a[env.env_start:env.env_start + E])
      torch.cuda.synchronize()
      mine = dict(step_type=bo.step_type.copy(), reward=bo.reward.copy(), success=bo.success.copy(),
                  frames=bo.frames.copy())
      everyone = [None] * world
      dist.all_gather_object(everyone, mine)
      want = {k: np.concatenate([p[k] for p in everyone]) for k in mine}
      ok = ok and ts.step_type.shape == (E_TOTAL,)
      ok = ok and np.array_equal(ts.step_type.cpu().numpy(), want['step_type'])
      ok = ok and np.allclose(ts.reward.cpu().numpy(), want['reward'], rtol=1e-13, atol=1e-12)
      ok = ok and np.array_equal(ts.success.cpu().numpy(), want['success'])
      ok = ok and np.array_equal(ts.observation['image'].cpu().numpy(), want['frames'])
      ok = ok and not ts.status.any().item()
      n_first += int((want['step_type'] == 0).sum())
      dist.barrier()   # nobody overwrites a frame slot a peer is still comparing
    ret[rank] = bool(ok) and n_first > 3 * E_TOTAL
    env.close()
  finally:
    dist.destroy_process_group()


def test_sharded_environment_two_ranks_matches_oracle():
  world = 2
  port = 29500 + (os.getpid() + 911) % 2000
  with mp.Manager() as m:
    ret = m.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
