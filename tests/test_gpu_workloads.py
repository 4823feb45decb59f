"""Parity ON THE BENCHMARKED WORKLOADS, at bench size ("untested configs").

For each BASELINE config C2..C5 the engine is built exactly as bench.py builds it
(workloads.build_engine, per-GPU env count), stepped max_episode_length + 5 times so that
every env auto-resets at least once, once through Engine.step (device buffers) and once
through swb_step_host (host buffers, frames rendered in env chunks), and a subset of >= 512
envs -- env 0, env E-1, both sides of every chunk boundary of swb_step_host, the rest random --
is compared with the CPU oracle stepping the same scenes and actions: positions, scene
cursors, step types, success flags and FRAMES bit-exact; rewards to the tolerances of
tests/test_gpu_parity.py.
"""
import os

import numpy as np
import pytest

from tests import fixtures

pytestmark = pytest.mark.gpu

N_SUBSET = 512


def _subset(E, rng):
  must = {0, E - 1}
  for c in range(1, 8):   # swb_step_host renders envs [E*c/8, E*(c+1)/8) per launch
    b = E * c // 8
    must.update((b - 1, b))
  rest = rng.choice(E, N_SUBSET, replace=False)
  idx = np.array(sorted(must | set(int(i) for i in rest)))[:max(N_SUBSET, len(must))]
  return np.array(sorted(set(idx.tolist()) | must))


def _threads():
  try:
    return max(1, min(16, len(os.sched_getaffinity(0))))
  except AttributeError:
    return 4


@pytest.mark.parametrize('host_call', [False, True], ids=['device', 'host'])
@pytest.mark.parametrize('key', ['c2', 'c3', 'c4', 'c5'])
def test_workload_matches_oracle_at_bench_size(key, host_call):
  import torch
  from spriteworld_b200 import workloads
  wl = workloads.WORKLOADS[key]()
  E = wl.n_envs
  T = wl.max_episode_length + 5
  K = T // wl.max_episode_length + 3
  eng, raster, scenes = workloads.build_engine(wl, E, K, device=0, seed=1000)
  actions = wl.sample_actions(np.random.RandomState(7), T, E)
  idx = _subset(E, np.random.RandomState(11))
  idx_t = torch.from_numpy(idx).to(eng.device)
  bo = fixtures.workload_oracle(wl, scenes, E, K, env_subset=idx)
  live = eng.state_tensors()
  frames = raster.new_frames()
  clustering = wl.nodes[-1]['kind'] == 'clustering'
  seen_first = np.zeros(len(idx), int)
  for t in range(T):
    if host_call:
      reward, step_type, success, status, fr = eng.step_host(actions[t], raster)
      reward, step_type, success, status, fr = (a[idx] for a in (reward, step_type, success, status, fr))
    else:
      res = eng.step(torch.from_numpy(actions[t]).to(eng.device), raster, frames)
      reward, step_type, success, status, fr = (
          a[idx_t].cpu().numpy() for a in (res.reward, res.step_type, res.success, res.status, res.frames))
    px, py = live['pos_x'][idx_t].cpu().numpy(), live['pos_y'][idx_t].cpu().numpy()
    cursor = live['cursor'][idx_t].cpu().numpy()
    fixtures.step_oracle_threads(bo, actions[t][idx], _threads())
    where = (key, 'host' if host_call else 'device', t)
    assert np.array_equal(step_type, bo.step_type), where
    assert np.array_equal(cursor, bo.cursor), where
    assert np.array_equal(success, bo.success), where
    assert np.array_equal(status, bo.err), where
    occupied = bo.cur['shape'] > 0
    assert np.array_equal(px[occupied], bo.cur['x'][occupied]), where
    assert np.array_equal(py[occupied], bo.cur['y'][occupied]), where
    mid = step_type != 0
    if clustering:
      np.testing.assert_allclose(reward[mid], bo.reward[mid], rtol=1e-6, atol=1e-9, err_msg=str(where))
    else:
      np.testing.assert_allclose(reward[mid], bo.reward[mid], rtol=1e-14, atol=1e-13,
                                 equal_nan=True, err_msg=str(where))
    diff = fr != bo.frames
    assert not diff.any(), where + (int(diff.sum()), idx[np.nonzero(diff.reshape(len(idx), -1).any(1))[0]][:8].tolist())
    seen_first += (step_type == 0)
  # every env of the subset went through at least one auto-reset after the initial one
  assert (seen_first >= 2).all(), (key, int((seen_first < 2).sum()))
  raster.close()
  eng.close()
