"""GPU tests of the reference-facing Python surface: the drop-in Environment, the
BatchedEnvironment, and the per-call plugin protocol (action_space.step / task.reward /
renderer.render on Python sprite lists), with expectations taken from the reference's own
tests (file:line cited) and from episodes stepped by the reference (tests/golden)."""
import importlib

import numpy as np
import pytest

from tests import fixtures

pytestmark = pytest.mark.gpu


def _sw():
  import spriteworld_b200.action_spaces as action_spaces
  import spriteworld_b200.environment as environment
  import spriteworld_b200.factor_distributions as distribs
  import spriteworld_b200.renderers as renderers
  import spriteworld_b200.sprite as sprite
  import spriteworld_b200.tasks as tasks
  return action_spaces, environment, distribs, renderers, sprite, tasks


# ---------------------------------------------------------------------------------------
# drop-in Environment against episodes stepped by the reference
# ---------------------------------------------------------------------------------------

DROPIN = [
    ('more_targets_f64', 'cobra.goal_finding_more_targets', 'test'),
    ('clustering', 'cobra.clustering', 'train'),
    ('sorting', 'cobra.sorting', 'train'),
    ('embodied', 'examples.goal_finding_embodied', None),
]


@pytest.mark.parametrize('fixture,config,mode', DROPIN)
def test_dropin_environment_reproduces_reference_episode(fixture, config, mode):
  """Same config, same seeds, same actions -> the TimeSteps the reference produced
  (make_golden.run_episodes: seed 1000+e before each env is built, envs stepped in order)."""
  _, environment, _, _, _, _ = _sw()
  ep = fixtures.Episodes(fixture)
  mod = importlib.import_module('spriteworld_b200.configs.' + config)
  envs = []
  for e in range(ep.E):
    np.random.seed(1000 + e)
    cfg = mod.get_config(mode) if mode else mod.get_config()
    envs.append(environment.Environment(**cfg))
  embodied = ep.meta['action']['kind'] == 'embodied'
  fe = ep.meta['frame_envs']
  for t in range(ep.T):
    for e, env in enumerate(envs):
      a = ep.actions[t, e]
      ts = env.step([int(a[0]), int(a[1])] if embodied else a)
      assert int(ts.step_type) == ep.step_type[t, e], (t, e)
      if ts.first():
        assert ts.reward is None and ts.discount is None
      else:
        tol = 1e-6 if fixture == 'clustering' else 1e-14
        np.testing.assert_allclose(ts.reward, ep.reward[t, e], rtol=tol, atol=1e-12)
        assert ts.discount == (0.0 if ts.last() else 1.0)
      n = len(env.state()['sprites'])
      got = np.array([[float(s.x), float(s.y)] for s in env.state()['sprites']])
      assert np.array_equal(got, ep.pos[t, e, ep.S - n:]), (t, e)
      assert bool(env.success()) == bool(ep.success[t, e])
      if e in fe:
        assert np.array_equal(ts.observation['image'], ep.frames[t, fe.index(e)]), (t, e)
  spec = envs[0].observation_spec()
  assert spec['image'].shape == (64, 64, 3) and spec['image'].dtype == np.uint8
  for env in envs:
    env.close()


def test_environment_cadence():
  """tests/environment_test.py:53-88: max_episode_length and termination -> FIRST cadence."""
  action_spaces, environment, distribs, renderers, _, tasks = _sw()
  from spriteworld_b200 import sprite_generators
  factors = distribs.Product([distribs.Continuous('x', 0.1, 0.9),
                              distribs.Continuous('y', 0.1, 0.9)])
  env = environment.Environment(
      task=tasks.NoReward(), action_space=action_spaces.SelectMove(), renderers={},
      init_sprites=sprite_generators.generate_sprites(factors, num_sprites=2),
      max_episode_length=7)
  assert env.step(env.action_space.sample()).first()
  for _ in range(3):
    for _ in range(6):
      ts = env.step(env.action_space.sample())
      assert ts.mid() and ts.reward == 0.0
    assert env.step(env.action_space.sample()).last()
    assert env.step(env.action_space.sample()).first()
  # a successful task ends the episode at once
  env2 = environment.Environment(
      task=tasks.FindGoalPosition(terminate_distance=10.), action_space=action_spaces.SelectMove(),
      renderers={'success': renderers.Success()},
      init_sprites=sprite_generators.generate_sprites(factors, num_sprites=2))
  assert env2.step(env2.action_space.sample()).first()
  ts = env2.step(env2.action_space.sample())
  assert ts.last() and ts.observation['success'] is True
  assert env2.step(env2.action_space.sample()).first()
  env.close(); env2.close()


# ---------------------------------------------------------------------------------------
# plugin protocol on Python sprite lists
# ---------------------------------------------------------------------------------------

def test_select_move_script():
  """tests/action_spaces_test.py:53-98."""
  action_spaces, _, _, _, sprite, _ = _sw()
  space = action_spaces.SelectMove(scale=0.5)
  sprites = [sprite.Sprite(x=0.55, y=0.5), sprite.Sprite(x=0.5, y=0.5)]
  script = [
      ([0.52, 0.52, 0.5, 0.48], False, [0.55, 0.5], [0.5, 0.49]),
      ([0.58, 0.5, 0.9, 0.9], False, [0.75, 0.7], [0.5, 0.49]),
      ([0.58, 0.5, 0.9, 0.9], False, [0.75, 0.7], [0.5, 0.49]),
      ([0.5, 0.5, 0.2, 0.5], False, [0.75, 0.7], [0.35, 0.49]),
      ([0.78, 0.74, 0.9, 0.9], False, [0.95, 0.9], [0.35, 0.49]),
      ([0.92, 0.9, 0.9, 0.5], True, [1., 0.9], [0.35, 0.49]),
      ([0.98, 0.9, 0.7, 0.9], False, [1.1, 1.1], [0.35, 0.49]),
  ]
  for action, keep, p0, p1 in script:
    space.step(np.array(action), sprites, keep_in_frame=keep)
    assert np.allclose(sprites[0].position, p0, atol=1e-5), action
    assert np.allclose(sprites[1].position, p1, atol=1e-5), action
  # motion cost (:41-51)
  cost = action_spaces.SelectMove(scale=1, motion_cost=1.).step(
      np.array([0.5, 0.5, 0.2, 0.75]), [], keep_in_frame=False)
  assert abs(cost - (-0.39)) < 0.01
  assert np.allclose(action_spaces.SelectMove(scale=0.5).get_motion(
      np.array([0.2, 0.5, 0.2, 0.75])), (-0.15, 0.125), atol=1e-4)


def test_drag_and_drop_script():
  """tests/action_spaces_test.py:123-168."""
  action_spaces, _, _, _, sprite, _ = _sw()
  space = action_spaces.DragAndDrop(scale=0.5)
  sprites = [sprite.Sprite(x=0.55, y=0.5), sprite.Sprite(x=0.5, y=0.5)]
  script = [
      ([0.52, 0.52, 0.52, 0.5], False, [0.55, 0.5], [0.5, 0.49]),
      ([0.58, 0.5, 0.98, 0.9], False, [0.75, 0.7], [0.5, 0.49]),
      ([0.58, 0.5, 0.9, 0.9], False, [0.75, 0.7], [0.5, 0.49]),
      ([0.5, 0.5, 0.2, 0.5], False, [0.75, 0.7], [0.35, 0.49]),
      ([0.78, 0.74, 0.98, 0.94], False, [0.85, 0.8], [0.35, 0.49]),
      ([0.82, 0.8, 1.3, 1.0], True, [1., 0.9], [0.35, 0.49]),
      ([0.99, 0.9, 1.19, 1.3], False, [1.1, 1.1], [0.35, 0.49]),
  ]
  for action, keep, p0, p1 in script:
    space.step(np.array(action), sprites, keep_in_frame=keep)
    assert np.allclose(sprites[0].position, p0, atol=1e-5), action
    assert np.allclose(sprites[1].position, p1, atol=1e-5), action


EMBODIED_CASES = [  # tests/action_spaces_test.py:185-241
    ([[0.5, 0.5], [0.2, 0.8]], (0, 0), [[0.5, 0.5], [0.2, 0.9]], True),
    ([[0.5, 0.5], [0.2, 0.8]], (1, 0), [[0.5, 0.5], [0.2, 0.9]], True),
    ([[0.5, 0.5], [0.45, 0.55]], (1, 3), [[0.6, 0.5], [0.55, 0.55]], True),
    ([[0.5, 0.5], [0.45, 0.55]], (1, 1), [[0.4, 0.5], [0.35, 0.55]], True),
    ([[0.5, 0.5], [0.45, 0.55]], (1, 2), [[0.5, 0.4], [0.45, 0.45]], True),
    ([[0.95, 0.02], [0.95, 0.05]], (1, 3), [[1., 0.02], [1., 0.05]], True),
    ([[0.95, 0.02], [0.95, 0.05]], (1, 3), [[1.05, 0.02], [1.05, 0.05]], False),
    ([[0.45, 0.55], [0.5, 0.5], [0.45, 0.55]], (1, 3), [[0.45, 0.55], [0.6, 0.5], [0.55, 0.55]],
     True),
]


@pytest.mark.parametrize('init,action,final,keep', EMBODIED_CASES)
def test_embodied_moves(init, action, final, keep):
  action_spaces, _, _, _, sprite, _ = _sw()
  space = action_spaces.Embodied(step_size=0.1)
  sprites = [sprite.Sprite(x=p[0], y=p[1], shape='square', scale=0.15) for p in init]
  space.step(action, sprites, keep_in_frame=keep)
  for s, p in zip(sprites, final):
    assert np.allclose(s.position, p, atol=1e-5)
  with pytest.raises(KeyError):
    space.step((0, 7), sprites, keep_in_frame=True)


def test_find_goal_tables():
  """tests/tasks_test.py:42-168 (real Sprites instead of mocks)."""
  _, _, distribs, _, sprite, tasks = _sw()

  def mk(positions):
    return [sprite.Sprite(x=p[0], y=p[1]) for p in positions]

  for pos, reward, success in [([[0., 0.]], -30.4, False), ([[0.4, 0.6]], -2.1, False),
                               ([[0.43, 0.56]], 0.4, True),
                               ([[0.48, 0.52], [0.4, 0.6]], 1.5, False),
                               ([[0.48, 0.52], [0.5, 0.5]], 8.6, True)]:
    task = tasks.FindGoalPosition(goal_position=(0.5, 0.5), terminate_distance=0.1)
    assert abs(task.reward(mk(pos)) - reward) < 0.1 and task.success(mk(pos)) == success
  for pos, w, reward, success in [([[0.43, 0.52]], (3, 1), -1.1, False),
                                  ([[0.3, 0.52]], (7, 2), -21.5, False),
                                  ([[0.3, 0.52]], (0.1, 0.2), 1.8, True)]:
    task = tasks.FindGoalPosition(goal_position=(0.5, 0.5), terminate_distance=0.1,
                                  weights_dimensions=w)
    assert abs(task.reward(mk(pos)) - reward) < 0.1 and task.success(mk(pos)) == success
  for pos, bonus, reward in [([[0.35, 0.52]], 1., 0.), ([[0.43, 0.52]], 3., 4.4),
                             ([[0.43, 0.52], [0.4, 0.55]], 1., 0.),
                             ([[0.43, 0.52], [0.43, 0.52]], 3., 5.7)]:
    task = tasks.FindGoalPosition(goal_position=(0.5, 0.5), terminate_distance=0.1,
                                  sparse_reward=True, terminate_bonus=bonus)
    assert abs(task.reward(mk(pos)) - reward) < 0.1
  sprites = [sprite.Sprite(x=0.45, y=0.45, c0=64), sprite.Sprite(x=0.45, y=0.55, c0=128),
             sprite.Sprite(x=0.55, y=0.45, c0=192), sprite.Sprite(x=0.4, y=0.4, c0=255)]
  for (lo, hi), reward, success in zip([(0, 65), (0, 129), (0, 193), (0, 256), (65, 256)],
                                       [1.5, 2.9, 4.4, 2.3, 0.9],
                                       [True, True, True, False, False]):
    task = tasks.FindGoalPosition(filter_distrib=distribs.Continuous('c0', lo, hi),
                                  goal_position=(0.5, 0.5), terminate_distance=0.1)
    assert abs(task.reward(sprites) - reward) < 0.1 and task.success(sprites) == success
  r = tasks.FindGoalPosition(filter_distrib=distribs.Continuous('c0', 0, 254),
                             goal_position=(0.5, 0.5), terminate_distance=0.1).reward(
                                 [sprite.Sprite(x=0.45, y=0.45, c0=255)])
  assert np.isnan(r)


def test_clustering_and_meta_tables():
  """tests/tasks_test.py:186-294, 339-409."""
  _, _, distribs, _, sprite, tasks = _sw()
  clusters = [distribs.Continuous('c0', 0, 129), distribs.Continuous('c0', 190, 256)]
  for positions, reward, success in [
      ([[0.2, 0.2], [0.21, 0.21], [0.8, 0.8], [0.81, 0.81]], 287.5, True),
      ([[0.2, 0.2], [0.25, 0.25], [0.8, 0.8], [0.81, 0.81]], 84.2, True),
      ([[0.2, 0.2], [0.53, 0.53], [0.8, 0.8], [0.81, 0.81]], 0.4, True),
      ([[0.2, 0.2], [0.53, 0.53], [0.8, 0.8], [0.9, 0.9]], -1.2, False)]:
    sprites = [sprite.Sprite(x=p[0], y=p[1], c0=c) for p, c in zip(positions, (64, 128, 192, 255))]
    task = tasks.Clustering(cluster_distribs=clusters)
    assert abs(task.reward(sprites) - reward) < 0.1 and task.success(sprites) == success
  six = [sprite.Sprite(x=0.2, y=0.2, c0=64), sprite.Sprite(x=0.3, y=0.3, c0=64),
         sprite.Sprite(x=0.8, y=0.9, c0=128), sprite.Sprite(x=0.9, y=0.8, c0=128),
         sprite.Sprite(x=0.8, y=0.9, c0=255), sprite.Sprite(x=0.9, y=0.8, c0=255)]
  three = [distribs.Continuous('c0', 0, 100), distribs.Continuous('c0', 100, 150),
           distribs.Continuous('c0', 200, 256)]
  assert abs(tasks.Clustering(cluster_distribs=three).reward(six) - 17.5) < 0.1
  # one populated cluster -> the reference raises sklearn's ValueError
  with pytest.raises(ValueError):
    tasks.Clustering(cluster_distribs=clusters).reward(
        [sprite.Sprite(x=0.1, y=0.1, c0=10), sprite.Sprite(x=0.2, y=0.2, c0=20)])
  # MetaAggregated over two goal tasks
  subtasks = [tasks.FindGoalPosition(filter_distrib=distribs.Continuous('c0', 0, 100),
                                     goal_position=(0.2, 0.2), terminate_distance=0.1),
              tasks.FindGoalPosition(filter_distrib=distribs.Continuous('c0', 200, 256),
                                     goal_position=(0.8, 0.8), terminate_distance=0.1)]
  sprites = [sprite.Sprite(x=0.2, y=0.25, c0=64), sprite.Sprite(x=0.8, y=0.7, c0=255)]
  r0, r1 = subtasks[0].reward(sprites), subtasks[1].reward(sprites)
  for agg, expect in (('sum', r0 + r1), ('max', max(r0, r1)), ('min', min(r0, r1)),
                      ('mean', (r0 + r1) / 2)):
    t = tasks.MetaAggregated(subtasks, reward_aggregator=agg, termination_criterion='any',
                             terminate_bonus=5.)
    assert t.success(sprites) is True
    assert abs(t.reward(sprites) - (expect + 5.)) < 1e-9
  t = tasks.MetaAggregated(subtasks, termination_criterion='all', terminate_bonus=5.)
  assert t.success(sprites) is False and abs(t.reward(sprites) - (r0 + r1)) < 1e-9


def test_pil_renderer_protocol():
  """tests/renderers/pil_renderer_test.py:45-88 through PILRenderer.render(sprites)."""
  _, _, _, renderers, sprite, _ = _sw()
  import colorsys

  def fixture():
    return [sprite.Sprite(x=0.75, y=0.95, shape='spoke_6', scale=0.2, c0=20, c1=50, c2=80),
            sprite.Sprite(x=0.2, y=0.3, shape='triangle', scale=0.1, c0=150, c1=255, c2=100),
            sprite.Sprite(x=0.7, y=0.5, shape='square', scale=0.3, c0=0, c1=255, c2=0),
            sprite.Sprite(x=0.5, y=0.5, shape='square', scale=0.3, c0=255, c1=0, c2=0)]

  image = renderers.PILRenderer(image_size=(64, 64), bg_color=(5, 6, 7)).render(fixture())
  assert list(image[5, 5]) == [5, 6, 7]
  image = renderers.PILRenderer(image_size=(64, 64)).render(fixture())
  assert list(image[32, 32]) == [255, 0, 0] and list(image[32, 50]) == [0, 255, 0]
  image = renderers.PILRenderer(image_size=(16, 16), anti_aliasing=5).render(fixture())
  assert list(image[4, 6]) == [0, 0, 0] and list(image[6, 6]) == [255, 0, 0]
  assert all(image[5, 6] >= [50, 0, 0]) and all(image[5, 6] <= [120, 30, 0])
  assert all(image[7, 6] >= [200, 0, 0]) and all(image[7, 6] <= [255, 50, 0])

  def to_rgb(c):
    return tuple((255 * np.array(colorsys.hsv_to_rgb(*c))).astype(np.uint8))

  s = sprite.Sprite(x=0.5, y=0.5, shape='square', c0=0.2, c1=0.5, c2=0.5)
  image = renderers.PILRenderer(image_size=(64, 64), color_to_rgb=to_rgb).render([s])
  assert list(image[32, 32]) == [114, 127, 63]
  # and the frames equal the ones the reference rendered
  shapes, cases = fixtures.render_cases()
  meta, arrs, frame = [c for c in cases if c[0]['name'] == 'fixture_aa5_64'][0]
  got = renderers.PILRenderer(image_size=(64, 64), anti_aliasing=5).render(fixture())
  assert np.array_equal(got, frame)


# ---------------------------------------------------------------------------------------
# BatchedEnvironment
# ---------------------------------------------------------------------------------------

def test_batched_environment_matches_oracle_with_pool_refill():
  """256 envs of the sorting config for 70 steps: the pool is refilled several times;
  every output must equal the CPU oracle stepping the same scenes."""
  _batched_environment_vs_oracle()


def test_batched_environment_refill_from_worker_processes():
  """The same with the refill's scenes sampled and packed by worker processes
  (_sampler_pool), three blocks of envs over two workers."""
  _batched_environment_vs_oracle(refill_threads=3, refill_procs=2)


def _batched_environment_vs_oracle(**env_kwargs):
  import torch
  from oracle import oracle
  from spriteworld_b200 import constants, environment, scene, sprite_generators
  from spriteworld_b200.configs.cobra import sorting
  cfg = sorting.get_config('train')
  E, K, T = 256, 4, 70
  uploads = []
  real_batch_of = sprite_generators.batch_of

  env = environment.BatchedEnvironment(n_envs=E, pool_depth=K, rng=np.random.RandomState(3),
                                       **dict(cfg, **env_kwargs))
  eng = env.engine
  # mirror of the device pool for the oracle, kept in sync by wrapping upload_scenes
  S = eng.n_slots
  pool = np.zeros((E, K, S), oracle.SPRITE_DTYPE)

  def record(batch, env_ids, ring_slots):
    for f in ('x', 'y', 'm00', 'm01', 'm10', 'm11', 'vx', 'vy', 'member', 'shape', 'pos_f32',
              'rgb'):
      pool[f][np.asarray(env_ids), np.asarray(ring_slots)] = batch[f]

  orig_upload = eng.upload_scenes

  def upload(batch, env_ids, ring_slots):
    record(batch, env_ids, ring_slots)
    return orig_upload(batch, env_ids, ring_slots)

  eng.upload_scenes = upload
  # the constructor already uploaded the first K scenes per env: re-read them
  env2_rng = np.random.RandomState(3)
  first = real_batch_of(cfg['init_sprites'], E * K, env2_rng)
  nodes, filters = cfg['task'].compile()
  b0 = scene.arrays_from_layout(first, S, filters, cfg['renderers']['image'].color_to_rgb)
  record(b0, np.repeat(np.arange(E), K), np.tile(np.arange(K), E))

  meta = dict(action=cfg['action_space'].compile(), keep_in_frame=True,
              max_episode_length=cfg['max_episode_length'], nodes=nodes)
  ocfg = fixtures.env_cfg_from_meta(meta)
  tab = oracle.shape_table(constants.SHAPES)
  rc = oracle.raster_cfg(64, 64, 5)
  bo = oracle.BatchOracle(ocfg, tab, rc, pool)
  bo.pool = pool            # share the live mirror (refills land in it)
  rng = np.random.RandomState(11)
  n_first = 0
  for t in range(T):
    a = rng.uniform(0, 1, (E, 4)).astype(np.float32)
    ts = env.step(a)
    bo.step(a)
    st = ts.step_type.cpu().numpy()
    assert np.array_equal(st, bo.step_type), t
    n_first += int((st == 0).sum())
    np.testing.assert_allclose(ts.reward.cpu().numpy(), bo.reward, rtol=1e-13, atol=1e-12)
    assert np.array_equal(ts.success.cpu().numpy(), bo.success)
    assert np.array_equal(ts.observation['image'].cpu().numpy(), bo.frames), t
    assert ts.discount.dtype == torch.float32
  assert n_first > E            # auto-resets happened beyond the initial one
  stats = env.refill_stats()
  assert stats['refills'] >= 1 and stats['scenes'] >= n_first - E
  env.close()


def test_batched_factor_observations_and_gym_surface():
  """SpriteFactors / Success in batched form and the gym-style wrappers."""
  import torch
  from spriteworld_b200 import environment, gym_wrapper, renderers
  from spriteworld_b200.configs.cobra import goal_finding_more_targets as cfgmod
  cfg = cfgmod.get_config('test')
  cfg['renderers'] = dict(cfg['renderers'], factors=renderers.SpriteFactors(),
                          success=renderers.Success())
  env = environment.BatchedEnvironment(n_envs=64, pool_depth=4, rng=np.random.RandomState(0), **cfg)
  venv = gym_wrapper.VectorGymWrapper(env)
  obs = venv.reset()
  assert obs['image'].shape == (64, 64, 64, 3) and obs['image'].dtype == torch.uint8
  f = obs['factors']
  assert f['mask'].all() and f['x'].shape == (64, 4)
  state = env.engine.download_state()
  assert np.array_equal(f['x'].cpu().numpy(), state['pos_x'].astype(np.float32))
  assert ((f['shape'] >= 1) & (f['shape'] <= 6)).all()
  assert torch.allclose(f['scale'], torch.full_like(f['scale'], 0.13))
  assert obs['success'].dtype == torch.bool
  a = torch.rand(64, 4, device=env.engine.device)
  obs, reward, done, info = venv.step(a)
  assert reward.shape == (64,) and done.dtype == torch.bool and not info['first'].any()
  venv.close()
  # single-env gym wrapper (reference gym_wrapper.py:91-109 semantics)
  env1 = environment.Environment(**cfgmod.get_config('test'))
  g = gym_wrapper.GymWrapper(env1)
  o = g.reset()
  assert o['image'].shape == (64, 64, 3)
  o, r, d, info = g.step(g.action_space.sample())
  assert isinstance(d, bool) and info['discount'] in (0.0, 1.0)
  assert np.array_equal(g.render(), o['image'])
  g.close()
