"""CPU tier: the host layer of the package (drop-in Environment, plugin protocol) driven
through the oracle-backed engine double (tests/oracle_engine.py).

The bodies are the GPU tests of tests/test_gpu_api.py, unchanged: the same expectations --
episodes stepped by the reference, tables from the reference's tests -- hold whether the
device or its CPU restatement executes the step, so a regression in the Python layer (scene
packing, RNG consumption order, the two-slot scene ring, TimeStep assembly, error mapping)
shows up here without a GPU.  Device arithmetic itself is not under test in this file.
"""
import pytest

from tests import oracle_engine
from tests import test_gpu_api as gpu_tests


@pytest.fixture(autouse=True)
def _oracle_engine(monkeypatch):
  oracle_engine.install(monkeypatch)


@pytest.mark.parametrize('fixture,config,mode', gpu_tests.DROPIN)
def test_dropin_environment_reproduces_reference_episode(fixture, config, mode):
  gpu_tests.test_dropin_environment_reproduces_reference_episode(fixture, config, mode)


def test_environment_cadence():
  gpu_tests.test_environment_cadence()


def test_select_move_script():
  gpu_tests.test_select_move_script()


def test_drag_and_drop_script():
  gpu_tests.test_drag_and_drop_script()


@pytest.mark.parametrize('init,action,final,keep', gpu_tests.EMBODIED_CASES)
def test_embodied_moves(init, action, final, keep):
  gpu_tests.test_embodied_moves(init, action, final, keep)


def test_find_goal_tables():
  gpu_tests.test_find_goal_tables()


def test_clustering_and_meta_tables():
  gpu_tests.test_clustering_and_meta_tables()


def test_pil_renderer_protocol():
  gpu_tests.test_pil_renderer_protocol()


def test_batched_environment_matches_oracle_with_pool_refill():
  gpu_tests.test_batched_environment_matches_oracle_with_pool_refill()


def test_batched_factor_observations_and_gym_surface():
  gpu_tests.test_batched_factor_observations_and_gym_surface()


def test_batched_environment_applies_action_noise():
  """SelectMove(noise_scale=...) (action_spaces.py:69-75) in the batched environment: same
  result as feeding pre-noised float64 actions drawn from an identically seeded stream."""
  import numpy as np
  from spriteworld_b200 import action_spaces, environment
  from spriteworld_b200.configs.cobra import goal_finding_more_targets as cfgmod

  def make(noise_scale, seed):
    cfg = cfgmod.get_config('train')
    cfg['action_space'] = action_spaces.SelectMove(scale=0.25, noise_scale=noise_scale)
    return environment.BatchedEnvironment(n_envs=16, pool_depth=6,
                                          rng=np.random.RandomState(seed), **cfg)

  noisy, plain = make(0.05, 3), make(None, 3)
  twin = np.random.RandomState(3)
  twin.set_state(plain._rng.get_state())     # the stream after both sampled their scene pools
  noisy.reset(); plain.reset()
  twin.normal(size=(16, 4))                  # reset() stepped once with a dummy action
  actions = np.random.RandomState(11).uniform(0, 1, (3, 16, 4)).astype(np.float32)
  for t in range(3):
    a = noisy.step(actions[t])
    pre = actions[t] + twin.normal(loc=0.0, scale=0.05, size=(16, 4))
    b = plain.step(pre)
    assert np.array_equal(noisy.engine.download_state()['pos_x'], plain.engine.download_state()['pos_x'])
    assert np.array_equal(a.reward.numpy(), b.reward.numpy())
    assert np.array_equal(a.observation['image'].numpy(), b.observation['image'].numpy())


def test_batched_environment_refill_never_serves_a_stale_scene():
  """The pool refill at the fastest consumption the protocol allows (max_episode_length=1:
  a scene every two steps).  Every scene an env lands on must have been uploaded for exactly
  that visit of its ring slot (never a leftover of the previous lap), and the outputs must
  equal the oracle stepping the same pool."""
  import numpy as np
  from oracle import oracle
  from spriteworld_b200 import constants, environment
  from spriteworld_b200.configs.cobra import sorting
  from tests import fixtures
  cfg = sorting.get_config('train')
  cfg['max_episode_length'] = 1           # FIRST, LAST, FIRST, ...: a scene every two steps
  E, K, T = 64, 5, 60
  env = environment.BatchedEnvironment(n_envs=E, pool_depth=K, rng=np.random.RandomState(5),
                                       **cfg)
  eng = env.engine
  S = eng.n_slots
  uploads = np.ones((E, K), np.int64)      # the constructor filled every slot once
  pushed = eng.upload_scenes

  def upload(batch, env_ids, ring_slots):
    np.add.at(uploads, (np.asarray(env_ids), np.asarray(ring_slots)), 1)
    return pushed(batch, env_ids, ring_slots)

  eng.upload_scenes = upload
  # an independent oracle over a live mirror of the pool (the double's own pool array)
  nodes, _ = cfg['task'].compile()
  ocfg = fixtures.env_cfg_from_meta(dict(action=cfg['action_space'].compile(), keep_in_frame=True,
                                         max_episode_length=1, nodes=nodes))
  bo = oracle.BatchOracle(ocfg, oracle.shape_table(constants.SHAPES), oracle.raster_cfg(64, 64, 5),
                          eng._bo.pool.copy())
  bo.pool = eng._bo.pool
  absolute = np.zeros(E, np.int64)         # index of the scene each env is on
  last_cursor = np.zeros(E, np.int64)
  rng = np.random.RandomState(11)
  for t in range(T):
    a = rng.uniform(0, 1, (E, 4)).astype(np.float32)
    ts = env.step(a)
    bo.step(a)
    cur = eng._bo.cursor.astype(np.int64)
    absolute += (cur - last_cursor) % K
    last_cursor = cur
    assert np.array_equal(uploads[np.arange(E), absolute % K], absolute // K + 1), t
    assert np.array_equal(ts.step_type.numpy(), bo.step_type), t
    assert np.array_equal(ts.observation['image'].numpy(), bo.frames), t
  assert absolute.min() >= 2 * K           # every ring went round at least twice
  env.close()


def test_batched_environment_async_refill_on_the_double():
  """The asynchronous refill (worker thread, snapshots taken ahead of need, underflow guard) with
  the oracle-backed engine double: at the fastest consumption the protocol allows no env ever
  lands on a scene that was not uploaded for exactly that visit of its ring slot, and the outputs
  equal an independent oracle's.  (On the device the same logic runs against streams and events:
  tests/test_gpu_api.py.)"""
  import numpy as np
  from oracle import oracle
  from spriteworld_b200 import constants, environment
  from spriteworld_b200.configs.cobra import sorting
  from tests import fixtures
  cfg = sorting.get_config('train')
  cfg['max_episode_length'] = 1           # FIRST, LAST, FIRST, ...: a scene every two steps
  E, K, T = 48, 6, 80
  env = environment.BatchedEnvironment(n_envs=E, pool_depth=K, rng=np.random.RandomState(5),
                                       refill='async', **cfg)
  eng = env.engine
  uploads = np.ones((E, K), np.int64)      # the constructor filled every slot once
  pushed = eng.upload_scenes

  def upload(batch, env_ids, ring_slots):
    np.add.at(uploads, (np.asarray(env_ids), np.asarray(ring_slots)), 1)
    return pushed(batch, env_ids, ring_slots)

  eng.upload_scenes = upload
  nodes, _ = cfg['task'].compile()
  ocfg = fixtures.env_cfg_from_meta(dict(action=cfg['action_space'].compile(), keep_in_frame=True,
                                         max_episode_length=1, nodes=nodes))
  bo = oracle.BatchOracle(ocfg, oracle.shape_table(constants.SHAPES), oracle.raster_cfg(64, 64, 5),
                          eng._bo.pool.copy())
  bo.pool = eng._bo.pool
  absolute = np.zeros(E, np.int64)
  last_cursor = np.zeros(E, np.int64)
  rng = np.random.RandomState(11)
  for t in range(T):
    if t in (30, 31, 32):                  # explicit resets count against the ring too
      ts = env.reset()
      bo.reset_next[:] = 1
      bo.step(np.zeros((E, 4), np.float32))
    else:
      a = rng.uniform(0, 1, (E, 4)).astype(np.float32)
      ts = env.step(a)
      bo.step(a)
    cur = eng._bo.cursor.astype(np.int64)
    absolute += (cur - last_cursor) % K
    last_cursor = cur
    assert np.array_equal(uploads[np.arange(E), absolute % K], absolute // K + 1), t
    assert np.array_equal(ts.step_type.numpy(), bo.step_type), t
    assert np.array_equal(ts.observation['image'].numpy(), bo.frames), t
  stats = env.refill_stats()
  assert stats['mode'] == 'async' and stats['refills'] >= 5
  assert absolute.min() >= 2 * K
  env.close()
