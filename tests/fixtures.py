"""Loads tests/golden/*.npz (generated from the real reference by make_golden.py) and
turns them into oracle records / configs.  Test infrastructure."""
import json
import os

import numpy as np

from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FIELDS = ('x', 'y', 'pos_f32', 'shape', 'angle', 'scale', 'c0', 'c1', 'c2', 'color_f32',
          'vx', 'vy', 'member', 'rgb')
SHAPE_NAMES = ('triangle', 'square', 'pentagon', 'hexagon', 'octagon', 'circle', 'star_4',
               'star_5', 'star_6', 'spoke_4', 'spoke_5', 'spoke_6')


def load(name):
  return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def shapes_of(blob):
  return {n: blob['shape_' + n] for n in SHAPE_NAMES}


def records_from_arrays(arrs):
  """dict of per-slot factor arrays (any leading shape) -> oracle SPRITE_DTYPE array."""
  shp = arrs['x'].shape
  rec = np.zeros(shp, oracle.SPRITE_DTYPE)
  rec['x'], rec['y'] = arrs['x'], arrs['y']
  rec['vx'], rec['vy'] = arrs['vx'], arrs['vy']
  rec['member'] = arrs['member']
  rec['shape'] = arrs['shape']
  rec['pos_f32'] = arrs['pos_f32']
  rec['rgb'] = arrs['rgb']
  flat = rec.reshape(-1)
  sc, an = arrs['scale'].reshape(-1), arrs['angle'].reshape(-1)
  for i in range(flat.shape[0]):
    if flat['shape'][i]:
      (flat['m00'][i], flat['m01'][i], flat['m10'][i],
       flat['m11'][i]) = oracle.centred_matrix(sc[i], an[i])
  return rec


def render_cases():
  blob = load('render_cases.npz')
  tab_shapes = shapes_of(blob)
  out = []
  for i in range(len(blob['names'])):
    meta = json.loads(str(blob['meta_%d' % i]))
    arrs = {f: blob['%s_%d' % (f, i)] for f in FIELDS}
    out.append((meta, arrs, blob['frame_%d' % i]))
  return tab_shapes, out


def env_cfg_from_meta(meta):
  cfg = oracle.EnvCfg()
  a = meta['action']
  cfg.action_kind = {'select_move': oracle.ACT_SELECT_MOVE,
                     'drag_and_drop': oracle.ACT_DRAG_AND_DROP,
                     'embodied': oracle.ACT_EMBODIED}[a['kind']]
  cfg.action_scale = a['scale']
  cfg.motion_cost = a['motion_cost']
  cfg.keep_in_frame = int(meta['keep_in_frame'])
  cfg.max_episode_length = meta['max_episode_length']
  cfg.n_nodes = len(meta['nodes'])
  for i, nd in enumerate(meta['nodes']):
    n = cfg.nodes[i]
    if nd['kind'] == 'find_goal':
      n.kind = oracle.TASK_FIND_GOAL
      n.filter_slot = nd['filter_slot']
      n.goal[0], n.goal[1] = nd['goal']
      n.weights[0], n.weights[1] = nd['weights']
      n.terminate_distance = nd['terminate_distance']
      n.terminate_bonus = nd['terminate_bonus']
      n.raw_reward_multiplier = nd['raw_reward_multiplier']
      n.sparse_reward = int(nd['sparse_reward'])
    elif nd['kind'] == 'clustering':
      n.kind = oracle.TASK_CLUSTERING
      n.n_clusters = len(nd['cluster_slots'])
      for j, s in enumerate(nd['cluster_slots']):
        n.cluster_slots[j] = s
      n.termination_threshold = nd['termination_threshold']
      n.terminate_bonus = nd['terminate_bonus']
      n.sparse_reward = int(nd['sparse_reward'])
      n.reward_range = nd['reward_range']
    elif nd['kind'] == 'meta':
      n.kind = oracle.TASK_META
      n.n_children = len(nd['children'])
      for j, c in enumerate(nd['children']):
        n.children[j] = c
      n.aggregator = oracle.AGG[nd['aggregator']]
      n.criterion = oracle.CRIT[nd['criterion']]
      n.terminate_bonus = nd['terminate_bonus']
    else:
      n.kind = oracle.TASK_NO_REWARD
  return cfg


class Episodes(object):
  """One episodes_<cfg>.npz fixture."""

  def __init__(self, name):
    blob = load('episodes_%s.npz' % name)
    self.meta = json.loads(str(blob['meta']))
    self.shapes = shapes_of(blob)
    self.actions = blob['actions']
    self.pos = blob['pos']
    self.reward = blob['reward']
    self.step_type = blob['step_type']
    self.success = blob['success']
    self.scene_idx = blob['scene_idx']
    self.frames = blob['frames']
    self.n_scenes = blob['n_scenes']
    self.scenes = {f: blob['scene_' + f] for f in FIELDS}
    self.T, self.E = self.reward.shape
    self.S = self.meta['n_slots']

  def oracle_parts(self, with_raster=True):
    cfg = env_cfg_from_meta(self.meta)
    tab = oracle.shape_table(self.shapes)
    rc = (oracle.raster_cfg(self.meta['width'], self.meta['height'], self.meta['aa'],
                            self.meta['bg']) if with_raster else None)
    pool = records_from_arrays(self.scenes)
    return cfg, tab, rc, pool


def workload_oracle(wl, scenes, n_envs, pool_depth, env_subset=None, max_episode_length=None):
  """BatchOracle over (a subset of the envs of) the scene pool a workload engine was built
  with: `scenes` is what workloads.build_engine returned, laid out (n_envs * pool_depth, S).
  Envs are independent, so an oracle over a subset of envs reproduces exactly those envs."""
  from spriteworld_b200 import constants
  idx = np.arange(n_envs) if env_subset is None else np.asarray(env_subset)
  rec = np.zeros((n_envs * pool_depth, wl.n_slots), oracle.SPRITE_DTYPE)
  for f in ('x', 'y', 'm00', 'm01', 'm10', 'm11', 'vx', 'vy', 'member', 'shape', 'pos_f32', 'rgb'):
    rec[f] = scenes[f]
  pool = rec.reshape(n_envs, pool_depth, wl.n_slots)[idx]
  cfg = env_cfg_from_meta(dict(
      action=wl.action, keep_in_frame=True,
      max_episode_length=max_episode_length or wl.max_episode_length,
      nodes=[dict(n, goal=list(n.get('goal', (0, 0))), weights=list(n.get('weights', (1, 1))))
             if n['kind'] == 'find_goal' else n for n in wl.nodes]))
  tab = oracle.shape_table(constants.SHAPES)
  rc = oracle.raster_cfg(wl.image_size[0], wl.image_size[1], wl.anti_aliasing)
  return oracle.BatchOracle(cfg, tab, rc, pool)


def step_oracle_threads(bo, actions, n_threads):
  """One BatchOracle.step split over threads (the C call releases the GIL)."""
  from concurrent.futures import ThreadPoolExecutor
  n_threads = max(1, min(n_threads, bo.E))
  edges = np.linspace(0, bo.E, n_threads + 1).astype(int)
  with ThreadPoolExecutor(n_threads) as ex:
    list(ex.map(lambda i: bo.step(actions, int(edges[i]), int(edges[i + 1])), range(n_threads)))
  return bo
