"""CPU tests of the host side: shape tables, factor distributions / sprite generators
(RNG-order compatibility with the reference, pinned by tests/golden/sampling.npz), task
compilation, scene arrays, the colour map, the C-ABI library's exported symbols."""
import ctypes
import importlib
import json
import os
import re

import numpy as np
import pytest

from spriteworld_b200 import _native, constants, scene
from spriteworld_b200 import factor_distributions as distribs
from spriteworld_b200 import sprite_generators as gen
from spriteworld_b200 import tasks
from spriteworld_b200.renderers import color_maps
from tests import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TYPE_CODES = {float: 0, np.float64: 1, np.float32: 2, int: 3, np.int32: 4, np.uint8: 5,
              np.int64: 6, str: 7}
FACTORS9 = ('x', 'y', 'angle', 'scale', 'c0', 'c1', 'c2', 'x_vel', 'y_vel')


def test_shape_tables_equal_reference():
  blob = fixtures.load('render_cases.npz')
  for name in constants.SHAPE_NAMES:
    assert np.array_equal(constants.SHAPES[name], blob['shape_' + name]), name
  assert [int(constants.ShapeType[n]) for n in constants.SHAPE_NAMES] == list(range(1, 13))


def _config_keys():
  blob = fixtures.load('sampling.npz')
  return sorted({k.rsplit('.', 1)[0] for k in blob.files if k.endswith('.count')})


@pytest.mark.parametrize('key', _config_keys())
def test_configs_sample_like_reference(key):
  """Same seed -> same sprites (values AND scalar types) as the reference's config."""
  blob = fixtures.load('sampling.npz')
  pkg, name, mode = key.split('.')
  mod = importlib.import_module('spriteworld_b200.configs.%s.%s' % (pkg, name))
  cfg = mod.get_config(mode) if mode != 'None' else mod.get_config()
  np.random.seed(int(blob['seed']))
  counts, values, types, shapes = [], [], [], []
  for _ in range(len(blob[key + '.count'])):
    sprites = cfg['init_sprites']()
    counts.append(len(sprites))
    for s in sprites:
      f = s.factors
      shapes.append(int(constants.ShapeType[f['shape']]))
      values.append([float(f[k]) for k in FACTORS9])
      types.append([TYPE_CODES[type(f[k])] for k in FACTORS9])
  assert counts == blob[key + '.count'].tolist()
  assert shapes == blob[key + '.shapes'].tolist()
  assert np.array_equal(np.array(values).reshape(-1, 9), blob[key + '.values'])
  assert np.array_equal(np.array(types, np.uint8).reshape(-1, 9), blob[key + '.types'])
  assert cfg['max_episode_length'] == int(blob[key + '.max_episode_length'])
  nodes, _ = cfg['task'].compile()
  assert nodes == json.loads(str(blob[key + '.task']))
  assert cfg['action_space'].compile() == json.loads(str(blob[key + '.action']))


def test_batch_sampling_supports_and_types():
  shared = distribs.Product([
      distribs.Continuous('x', 0.1, 0.9), distribs.Continuous('y', 0.1, 0.9),
      distribs.Discrete('shape', ['square', 'triangle', 'circle']),
      distribs.Discrete('scale', [0.13]), distribs.Continuous('c1', 0.3, 1.),
      distribs.Continuous('c2', 0.9, 1.)])
  hue = distribs.Continuous('c0', 0., 0.4)
  other = distribs.Continuous('c0', 0.5, 0.9)
  g = gen.chain_generators(
      gen.shuffle(gen.chain_generators(
          gen.generate_sprites(distribs.Product([hue, shared]), 2),
          gen.generate_sprites(distribs.Product([other, shared]),
                               lambda: np.random.randint(1, 4)))),
      gen.generate_sprites(distribs.Product([
          distribs.Continuous('x', .1, .9), distribs.Continuous('y', .1, .9),
          distribs.Discrete('shape', ['circle']), distribs.Discrete('scale', [0.07]),
          distribs.Discrete('c0', [1.]), distribs.Discrete('c1', [0.]),
          distribs.Discrete('c2', [1.])]), 1))
  rng = np.random.RandomState(3)
  lay = g.batch(500, rng)
  b = scene.arrays_from_layout(lay, 6, [hue], color_maps.hsv_to_rgb)
  occupied = b['shape'] > 0
  assert np.array_equal(occupied.sum(1), lay.count)
  assert (lay.count >= 4).all() and (lay.count <= 6).all()
  assert occupied[:, -1].all() and (b['shape'][:, -1] == 6).all()          # body is last
  assert (b['rgb'][:, -1] == 255).all() and (b['pos_f32'][:, -1] == 1).all()
  for j in range(6):                                                        # padded at the front
    assert (occupied[:, j] <= occupied[:, min(j + 1, 5)]).all()
  assert (b['member'][occupied].astype(bool).sum() == 2 * 500)
  assert ((b['x'][occupied] >= 0.1) & (b['x'][occupied] <= 0.9)).all()
  m = scene.transform_matrix(0.13, 0)
  assert np.allclose(b['m00'][:, 0][occupied[:, 0]], m[0])
  # the scalar path through Sprite objects gives the same arrays as the batch path
  np.random.seed(4)
  sprites = [g() for _ in range(40)]
  b2 = scene.arrays_from_layout(gen.layout_from_sprite_lists(sprites), 6, [hue],
                                color_maps.hsv_to_rgb)
  for sc, sp in enumerate(sprites):
    for j, s in enumerate(sp):
      k = 6 - len(sp) + j
      assert b2['x'][sc, k] == float(s.x) and b2['shape'][sc, k] == int(
          constants.ShapeType[s.shape])
      assert tuple(b2['rgb'][sc, k]) == tuple(int(v) for v in color_maps.hsv_to_rgb(s.color))
      assert b2['member'][sc, k] == int(hue.contains(s.factors))
      assert (b2['m00'][sc, k], b2['m01'][sc, k], b2['m10'][sc, k],
              b2['m11'][sc, k]) == s.transform


def test_distribution_algebra():
  a = distribs.Continuous('x', 0., 1.)
  b = distribs.Continuous('x', 0.5, 2.)
  inter = distribs.Intersection([a, b])
  minus = distribs.SetMinus(a, b)
  sel = distribs.Selection(distribs.Product([a, distribs.Discrete('s', ['p', 'q'])]),
                           distribs.Discrete('s', ['q']))
  mix = distribs.Mixture([a, b], probs=[0.2, 0.8])
  rng = np.random.RandomState(0)
  for d, check in ((inter, lambda v: 0.5 <= v < 1), (minus, lambda v: 0 <= v < 0.5),
                   (mix, lambda v: 0 <= v < 2)):
    for _ in range(50):
      s = d.sample(rng)
      assert check(s['x']) and d.contains(s)
    cols = d.sample_batch(300, rng)
    assert all(check(v) for v in cols['x']) and d.contains_batch(cols).all()
  cols = sel.sample_batch(100, rng)
  assert (cols['s'] == 'q').all() and sel.contains_batch(cols).all()
  with pytest.raises(ValueError):
    distribs.Product([a, b])
  with pytest.raises(ValueError):
    distribs.Mixture([a, distribs.Continuous('y', 0, 1)])
  with pytest.raises(KeyError):
    a.contains({'y': 0.5})
  with pytest.raises(ValueError):
    distribs.Intersection([a, distribs.Continuous('x', 5., 6.)]).sample(rng)
  assert 'Continuous' in str(inter) and str(mix).count('Continuous') == 2


def test_task_compile_rejects_position_filters():
  t = tasks.FindGoalPosition(filter_distrib=distribs.Continuous('x', 0, 0.5))
  with pytest.raises(NotImplementedError):
    t._filters_static()
  with pytest.raises(ValueError):
    tasks.MetaAggregated([tasks.NoReward()], reward_aggregator='median')


def test_hsv_batch_matches_scalar():
  rng = np.random.RandomState(1)
  c = rng.uniform(0, 1, (3000, 3)).astype(np.float32)
  c[:20, 1] = 0
  ref = np.array([color_maps.hsv_to_rgb(tuple(r)) for r in c])
  assert np.array_equal(color_maps.hsv_to_rgb_batch(c[:, 0], c[:, 1], c[:, 2], True), ref)
  cd = rng.uniform(0, 1, (3000, 3))
  refd = np.array([color_maps.hsv_to_rgb(tuple(float(v) for v in r)) for r in cd])
  assert np.array_equal(color_maps.hsv_to_rgb_batch(cd[:, 0], cd[:, 1], cd[:, 2], False), refd)


def test_cabi_library_exports_every_declared_symbol():
  """The shared library loads (no GPU needed) and exports what include/*.h declares."""
  header = open(os.path.join(ROOT, 'include', 'spriteworld_b200.h')).read()
  declared = set(re.findall(r'\b(swb_[a-z_]+)\s*\(', header))
  assert len(declared) >= 18
  lib = ctypes.CDLL(_native.lib_path())
  for name in declared:
    assert hasattr(lib, name), name
  assert declared == set(_native.EXPORTS)
  L = _native.load()
  assert L.swb_sizeof_config() == ctypes.sizeof(_native.Config)
  assert L.swb_version() >= 1
  # error path without a GPU: null arguments are rejected with a message, not a crash
  assert L.swb_engine_create(None, None) != 0
  assert b'null' in L.swb_last_error()


def test_product_has_no_cpu_fallback():
  """The engine refuses to run without CUDA instead of silently computing on the host."""
  import torch
  if torch.cuda.is_available():
    pytest.skip('CUDA present')
  from spriteworld_b200 import engine
  with pytest.raises(_native.NativeError):
    engine.Engine(1, 1, 1, dict(kind='select_move', scale=1.0), [dict(kind='no_reward')],
                  constants.SHAPES)
  # and nothing under spriteworld_b200/ imports the oracle
  for dirpath, _, files in os.walk(os.path.join(ROOT, 'spriteworld_b200')):
    for f in files:
      if f.endswith('.py'):
        src = open(os.path.join(dirpath, f)).read()
        assert 'import oracle' not in src and 'from oracle' not in src, f


def test_generator_sprite_bounds():
  """SpriteGenerator.max_sprites: exact for fixed counts, None when a user callable draws
  the count (the batched environment sizes its sprite slots from it)."""
  import numpy as np
  from spriteworld_b200 import factor_distributions as distribs
  from spriteworld_b200 import sprite_generators as gen
  f = distribs.Product([distribs.Continuous('x', 0., 1.), distribs.Continuous('y', 0., 1.)])
  two, three = gen.generate_sprites(f, 2), gen.generate_sprites(f, num_sprites=3)
  drawn = gen.generate_sprites(f, num_sprites=lambda: np.random.randint(1, 4))
  assert (two.max_sprites, three.max_sprites, drawn.max_sprites) == (2, 3, None)
  assert gen.chain_generators(two, three).max_sprites == 5
  assert gen.sample_generator([two, three]).max_sprites == 3
  assert gen.shuffle(gen.chain_generators(two, three)).max_sprites == 5
  assert gen.chain_generators(two, drawn).max_sprites is None
  assert gen.shuffle(drawn).max_sprites is None


def test_sampler_pool_draws_what_the_in_process_sampler_draws():
  """_sampler_pool workers (spawned processes, results through shared memory) return exactly
  the arrays scene.arrays_from_layout gives in this process for the same seed, also for a
  request larger than the shared block (pickled through the pipe)."""
  import numpy as np
  from spriteworld_b200 import _sampler_pool, scene, sprite_generators
  from spriteworld_b200.configs.cobra import sorting
  cfg = sorting.get_config('train')
  _, filters = cfg['task'].compile()
  color_to_rgb = cfg['renderers']['image'].color_to_rgb
  S = cfg['init_sprites'].max_sprites
  pool = _sampler_pool.SamplerPool(2, cfg['init_sprites'], S, filters, color_to_rgb, capacity=64)
  try:
    for worker, n, seed in ((0, 40, 11), (1, 64, 12), (1, 100, 13)):
      got = {k: np.array(v) for k, v in pool.sample(worker, n, seed).items()}
      np.random.seed(seed ^ 0x5BD1E995)
      layout = sprite_generators.batch_of(cfg['init_sprites'], n, np.random.RandomState(seed))
      want = scene.arrays_from_layout(layout, S, filters, color_to_rgb)
      assert sorted(got) == sorted(want)
      for k in want:
        assert got[k].dtype == want[k].dtype and np.array_equal(got[k], want[k]), (k, n)
  finally:
    pool.close()


@pytest.mark.parametrize('name', ['cobra.exploration', 'cobra.sorting', 'cobra.clustering',
                                  'examples.goal_finding_clustering', 'examples.goal_finding_embodied'])
def test_batched_packing_equals_packing_the_sprites_one_scene_at_a_time(name):
  """scene.arrays_from_layout on a batch layout (several typed tables, merged when their columns
  agree, ragged sprite counts, object columns) against the plain route: build the Sprite objects
  of every scene from the same sampled factors and pack each scene on its own."""
  import importlib
  from spriteworld_b200 import scene, sprite, sprite_generators
  cfg = importlib.import_module('spriteworld_b200.configs.' + name).get_config('train')
  _, filters = cfg['task'].compile()
  color_to_rgb = cfg['renderers']['image'].color_to_rgb
  np.random.seed(4)
  layout = sprite_generators.batch_of(cfg['init_sprites'], 96, np.random.RandomState(9))
  n_slots = max(1, int(layout.count.max()))
  got = scene.arrays_from_layout(layout, n_slots, filters, color_to_rgb)
  for i in range(layout.n):
    sprites = []
    for j in range(int(layout.count[i])):
      cols = layout.tables[layout.ref_table[i, j]].columns
      row = layout.ref_row[i, j]
      sprites.append(sprite.Sprite(**{k: v[row] for k, v in cols.items() if not k.startswith('_')}))
    one = scene.arrays_from_layout(sprite_generators.layout_from_sprite_lists([sprites]), n_slots,
                                   filters, color_to_rgb)
    for k in one:
      assert np.array_equal(got[k][i], one[k][0]), (name, i, k)


def _pack_both_ways(layout, n_slots, filters, color_to_rgb, monkeypatch):
  from spriteworld_b200 import _host_pack, scene
  native = scene.arrays_from_layout(layout, n_slots, filters, color_to_rgb)
  with monkeypatch.context() as m:
    m.setattr(_host_pack, 'pack', lambda *a, **k: False)
    plain = scene.arrays_from_layout(layout, n_slots, filters, color_to_rgb)
  return native, plain


def test_native_packing_equals_the_numpy_path(monkeypatch):
  """csrc/swb_host_pack.c (one pass per sprite: gathers, HSV colour map in the factors' float
  type, shape ids by object pointer) against the NumPy path it shortcuts, value for value: the
  shipped configs, float64 and integer-RGB colours, grey and hue-1.0 corner cases, ragged counts."""
  import importlib
  from spriteworld_b200 import _host_pack, factor_distributions as distribs
  from spriteworld_b200 import sprite_generators as gen
  from spriteworld_b200.renderers import color_maps
  if _host_pack.load() is None:
    pytest.skip('libswb_host.so has not been built')
  used = []
  real = _host_pack.pack
  monkeypatch.setattr(_host_pack, 'pack', lambda *a, **k: used.append(real(*a, **k)) or used[-1])
  cases = []
  for name in ('cobra.sorting', 'cobra.clustering', 'cobra.exploration', 'cobra.goal_finding_more_targets',
               'examples.goal_finding_clustering'):
    cfg = importlib.import_module('spriteworld_b200.configs.' + name).get_config('train')
    cases.append((name, cfg['init_sprites'], cfg['task'].compile()[1], cfg['renderers']['image'].color_to_rgb))
  corner = distribs.Product([
      distribs.Continuous('x', 0., 1.), distribs.Continuous('y', 0., 1., dtype='float64'),
      distribs.Discrete('shape', ['star_5', 'circle', 'spoke_4']), distribs.Continuous('angle', 0., 360.),
      distribs.Continuous('scale', 0.05, 0.3, dtype='float64'),
      distribs.Discrete('c0', [0.0, 1.0, 0.999999, 0.16666667, 0.5]), distribs.Discrete('c1', [0.0, 1.0, 0.3]),
      distribs.Continuous('c2', 0., 1., dtype='float64'), distribs.Continuous('x_vel', -.1, .1)])
  cases.append(('float64 hsv corners', gen.generate_sprites(corner, lambda: np.random.randint(0, 5)),
                [distribs.Continuous('c2', 0.2, 0.7)], color_maps.hsv_to_rgb))
  rgb = distribs.Product([distribs.Continuous('x', 0., 1.), distribs.Continuous('y', 0., 1.),
                          distribs.Discrete('c0', [0, 128, 255]), distribs.Discrete('c1', [3, 200]),
                          distribs.Discrete('c2', [77])])
  cases.append(('integer rgb', gen.generate_sprites(rgb, 3), [], None))
  for name, generator, filters, color_to_rgb in cases:
    np.random.seed(2)
    layout = gen.batch_of(generator, 300, np.random.RandomState(1))
    n_slots = max(1, int(layout.count.max()))
    before = len(used)
    native, plain = _pack_both_ways(layout, n_slots, filters, color_to_rgb, monkeypatch)
    if name != 'examples.goal_finding_clustering':   # (its tables have different columns: NumPy path)
      assert used[before:] == [True], (name, used[before:])   # the native pass did run
    for k in plain:
      assert native[k].dtype == plain[k].dtype and np.array_equal(native[k], plain[k]), (name, k)
