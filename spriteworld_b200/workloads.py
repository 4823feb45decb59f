"""Synthetic workloads of BASELINE.json's configs.

Each workload fixes the action space, task tree, renderer and a vectorised scene sampler
that draws from the same factor distributions as the reference config it is modelled on
(file:line in each docstring), with the enlarged sprite counts BASELINE.json names.
Used by bench.py and the scale tests; the configs/ package is the drop-in surface for
the shipped reference configs.
"""
import numpy as np

from spriteworld_b200 import constants, scene
from spriteworld_b200.renderers import color_maps

SQUARE, TRIANGLE, CIRCLE = (int(constants.ShapeType[n]) for n in ('square', 'triangle', 'circle'))


def _common_factors(rng, n, s):
  """x, y in U[0.1, 0.9) float32; shape in {square, triangle, circle}; scale 0.13; c1 in
  [0.3, 1), c2 in [0.9, 1) float32 (e.g. goal_finding_more_targets.py:54-61)."""
  f32 = lambda a: a.astype(np.float32).astype(np.float64)
  return dict(
      x=f32(rng.uniform(0.1, 0.9, (n, s))), y=f32(rng.uniform(0.1, 0.9, (n, s))),
      shape=np.array([SQUARE, TRIANGLE, CIRCLE], np.uint8)[rng.randint(0, 3, (n, s))],
      scale=np.full((n, s), 0.13), angle=np.zeros((n, s)),
      c1=f32(rng.uniform(0.3, 1.0, (n, s))), c2=f32(rng.uniform(0.9, 1.0, (n, s))))


def _hue(rng, bands, n, s):
  """bands: (s, 2) or (n, s, 2) [lo, hi) hue ranges; float32 samples."""
  bands = np.broadcast_to(np.asarray(bands, np.float64), (n, s, 2))
  h = rng.uniform(bands[..., 0], bands[..., 1])
  return h.astype(np.float32).astype(np.float64)


def _shuffle_slots(rng, arrs, lo=0, hi=None):
  """Per-scene random z-order of slots [lo, hi) (sprite_generators.shuffle :101-128)."""
  n, s = arrs['x'].shape
  hi = s if hi is None else hi
  order = np.argsort(rng.uniform(size=(n, hi - lo)), axis=1) + lo
  rows = np.arange(n)[:, None]
  for k, v in arrs.items():
    v[:, lo:hi] = v[rows, order]
  return arrs


def _finish(arrs, member, color_f32=True):
  n, s = arrs['x'].shape
  rgb = color_maps.hsv_to_rgb_batch(arrs['c0'], arrs['c1'], arrs['c2'],
                                    np.broadcast_to(color_f32, (n, s)))
  zeros = np.zeros((n, s))
  return scene.batch_from_factor_arrays(
      arrs['x'], arrs['y'], np.ones((n, s), np.uint8), arrs['shape'], arrs['angle'],
      arrs['scale'], arrs['c0'], arrs['c1'], arrs['c2'], zeros, zeros, member, rgb)


class Workload(object):
  name = None
  n_envs = None       # per GPU
  n_slots = None
  image_size = (64, 64)
  anti_aliasing = 5
  max_episode_length = None
  action = None
  nodes = None
  bytes_per_env_step = None   # SURVEY.md 8(d) algorithmic bytes

  def sample_scenes(self, rng, n):
    raise NotImplementedError

  def sample_actions(self, rng, steps, n_envs):
    if self.action['kind'] == 'embodied':
      return np.stack([rng.randint(0, 2, (steps, n_envs)), rng.randint(0, 4, (steps, n_envs))],
                      -1).astype(np.int32)
    return rng.uniform(0, 1, (steps, n_envs, 4)).astype(np.float32)

  def algorithmic_bytes(self):
    w, h = self.image_size
    a = 2 if self.action['kind'] == 'embodied' else 16
    return h * w * 3 + self.n_slots * 40 + self.n_slots * 8 + a + 8 + 2

  def plugin_config(self):
    """The workload as the reference's config dict (task, action_space, renderers,
    init_sprites, max_episode_length) built from this package's plugin classes, i.e. what a
    user of the reference would write; drives BatchedEnvironment in bench.py's `api` number."""
    raise NotImplementedError('%s has no plugin-API form yet' % self.name)


class GoalFinding(Workload):
  """C2: goal_finding SelectMove, 4096 envs x 5 sprites, 64x64
  (configs/cobra/goal_finding_more_targets.py:54-86, common.py:26-38)."""
  name = 'goal_finding_select_move_4096x5_64x64'
  n_envs, n_slots, max_episode_length = 4096, 5, 20
  action = dict(kind='select_move', scale=0.25, motion_cost=0.0)
  nodes = [dict(kind='find_goal', filter_slot=0, goal=(0.5, 0.5), weights=(1, 1),
                terminate_distance=0.075, terminate_bonus=0.0, raw_reward_multiplier=50,
                sparse_reward=False)]

  def plugin_config(self):
    from spriteworld_b200 import factor_distributions as distribs
    from spriteworld_b200 import sprite_generators as gen
    from spriteworld_b200 import tasks
    from spriteworld_b200.configs.cobra import common
    shared = distribs.Product(common.body_factors())
    target_hue = distribs.Continuous('c0', 0., 0.4)
    distractor_hue = distribs.Continuous('c0', 0.5, 0.9)
    sprite_gen = gen.shuffle(gen.chain_generators(
        gen.generate_sprites(distribs.Product([target_hue, shared]), num_sprites=2),
        gen.generate_sprites(distribs.Product([distractor_hue, shared]), num_sprites=3)))
    task = tasks.FindGoalPosition(filter_distrib=target_hue, terminate_distance=0.075)
    cfg = common.config(task, sprite_gen, self.max_episode_length, __file__, None)
    cfg.pop('metadata')
    return cfg

  def sample_scenes(self, rng, n):
    s = self.n_slots
    a = _common_factors(rng, n, s)
    bands = [(0., 0.4)] * 2 + [(0.5, 0.9)] * 3   # 2 targets + 3 distractors
    a['c0'] = _hue(rng, bands, n, s)
    _shuffle_slots(rng, a)
    member = ((a['c0'] >= 0.) & (a['c0'] < 0.4)).astype(np.uint32)
    return _finish(a, member)


class Clustering(Workload):
  """C3: clustering SelectMove, 16384 envs x 9 sprites, 64x64
  (configs/cobra/clustering.py:41-97): blue/green hue clusters split 5/4."""
  name = 'clustering_select_move_16384x9_64x64'
  n_envs, n_slots, max_episode_length = 16384, 9, 50
  action = dict(kind='select_move', scale=0.25, motion_cost=0.0)
  nodes = [dict(kind='clustering', cluster_slots=[0, 1], termination_threshold=2.5,
                terminate_bonus=0.0, sparse_reward=False, reward_range=10.0)]

  def sample_scenes(self, rng, n):
    s = self.n_slots
    a = _common_factors(rng, n, s)
    bands = [(0.55, 0.65)] * 5 + [(0.27, 0.37)] * 4
    a['c0'] = _hue(rng, bands, n, s)
    _shuffle_slots(rng, a)
    c0 = a['c0']
    member = (((c0 >= 0.55) & (c0 < 0.65)).astype(np.uint32) |
              (((c0 >= 0.27) & (c0 < 0.37)).astype(np.uint32) << 1))
    return _finish(a, member)


SORT_BANDS = ((0.9, 1.0), (0.55, 0.65), (0.27, 0.37), (0.73, 0.83), (0.1, 0.2))
SORT_GOALS = ((0.75, 0.75), (0.75, 0.25), (0.25, 0.75), (0.25, 0.25), (0.5, 0.5))


class Sorting(Workload):
  """C4: sorting (MetaAggregated goal_finding), 65536 envs x 6 sprites over 8 GPUs
  (configs/cobra/sorting.py:40-67, 84-123): one sprite per hue band + one from a random band."""
  name = 'sorting_select_move_65536x6_64x64'
  n_envs, n_slots, max_episode_length = 8192, 6, 50
  action = dict(kind='select_move', scale=0.25, motion_cost=0.0)
  nodes = [dict(kind='find_goal', filter_slot=i, goal=SORT_GOALS[i], weights=(1, 1),
                terminate_distance=0.075, terminate_bonus=0.0, raw_reward_multiplier=20.,
                sparse_reward=False) for i in range(5)] + [
                    dict(kind='meta', children=[0, 1, 2, 3, 4], aggregator='sum',
                         criterion='all', terminate_bonus=0.0)]

  def sample_scenes(self, rng, n):
    s = self.n_slots
    a = _common_factors(rng, n, s)
    bands = np.empty((n, s, 2))
    bands[:, :5] = np.asarray(SORT_BANDS)[None]
    bands[:, 5] = np.asarray(SORT_BANDS)[rng.randint(0, 5, n)]
    a['c0'] = _hue(rng, bands, n, s)
    _shuffle_slots(rng, a)
    member = np.zeros((n, s), np.uint32)
    for i, (lo, hi) in enumerate(SORT_BANDS):
      member |= ((a['c0'] >= lo) & (a['c0'] < hi)).astype(np.uint32) << i
    return _finish(a, member)


class Embodied(Workload):
  """C5: goal_finding_embodied, 32768 envs x 8 sprites, 128x128 over 8 GPUs
  (configs/examples/goal_finding_embodied.py:53-111): 3 targets + 4 distractors shuffled,
  body (circle, scale 0.07, HSV (1, 0, 1) as Python floats) in the last slot."""
  name = 'goal_finding_embodied_32768x8_128x128'
  n_envs, n_slots, max_episode_length = 4096, 8, 50
  image_size = (128, 128)
  action = dict(kind='embodied', scale=0.05, motion_cost=0.0)
  nodes = GoalFinding.nodes

  def sample_scenes(self, rng, n):
    s = self.n_slots
    a = _common_factors(rng, n, s)
    bands = [(0., 0.4)] * 3 + [(0.5, 0.9)] * 4 + [(1.0, 1.0)]
    a['c0'] = _hue(rng, bands, n, s)
    _shuffle_slots(rng, a, 0, s - 1)
    a['shape'][:, -1] = CIRCLE
    a['scale'][:, -1] = 0.07
    a['c0'][:, -1], a['c1'][:, -1], a['c2'][:, -1] = 1.0, 0.0, 1.0
    member = ((a['c0'] >= 0.) & (a['c0'] < 0.4)).astype(np.uint32)
    color_f32 = np.ones((n, s), bool)
    color_f32[:, -1] = False
    return _finish(a, member, color_f32)


WORKLOADS = {'c2': GoalFinding, 'c3': Clustering, 'c4': Sorting, 'c5': Embodied}


def build_engine(wl, n_envs, pool_depth, device=0, seed=1000, max_episode_length=None):
  """Engine + raster + uploaded scene pool for workload `wl`.  Returns (engine, raster, scenes)."""
  from spriteworld_b200 import engine as engine_lib
  eng = engine_lib.Engine(
      n_envs, wl.n_slots, pool_depth, wl.action, wl.nodes, constants.SHAPES, keep_in_frame=True,
      max_episode_length=max_episode_length or wl.max_episode_length, device=device)
  rng = np.random.RandomState(seed)
  scenes = wl.sample_scenes(rng, n_envs * pool_depth)
  eng.upload_scenes(scenes, np.repeat(np.arange(n_envs), pool_depth),
                    np.tile(np.arange(pool_depth), n_envs))
  raster = engine_lib.Raster(eng, wl.image_size[0], wl.image_size[1], wl.anti_aliasing)
  return eng, raster, scenes
