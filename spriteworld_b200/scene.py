"""Host-side scene batches: what `init_sprites()` returned, reduced to the struct-of-arrays
the engine uploads (include/spriteworld_b200.h: swb_scene_soa).

A scene batch is a dict of numpy arrays shaped (n_scenes, n_slots[, ...]):
  x, y            float64   initial position (float32-valued where pos_f32)
  m00..m11        float64   centred-path transform, scale then rotate (sprite.py:96-101)
  vx, vy          float64   velocity
  member          uint32    bit i = task filter i contains the sprite (tasks.py:136,201)
  shape           uint8     ShapeType id, 0 = empty slot (slots are padded at the FRONT)
  pos_f32         uint8     the reference's position ndarray would be float32
  rgb             uint8 x3  color_to_rgb(sprite.color) (pil_renderer.py:82)
  factors         float32x5 scale, angle, c0, c1, c2 (for factor observations)
"""
import math

import numpy as np

F64_FIELDS = ('x', 'y', 'm00', 'm01', 'm10', 'm11', 'vx', 'vy')


def transform_matrix(scale, angle):
  """Matrix of Affine2D().scale(s) + Affine2D().rotate_deg(a) (sprite.py:96-101).

  rotate_deg goes through math.radians / math.cos / math.sin (libm), and the composite
  matrix is rotate . scale, whose entries are single products.
  """
  s = float(scale)
  th = math.radians(angle)
  c, sn = math.cos(th), math.sin(th)
  return c * s, (-sn) * s, sn * s, c * s


_matrix_cache = {}


def transform_matrices(scale, angle):
  """Vectorised transform_matrix over arrays (values are cached: configs use few)."""
  scale = np.asarray(scale, dtype=np.float64)
  angle = np.asarray(angle, dtype=np.float64)
  out = np.empty(scale.shape + (4,), np.float64)
  flat_s, flat_a, flat_o = scale.reshape(-1), angle.reshape(-1), out.reshape(-1, 4)
  for i in range(flat_s.shape[0]):
    key = (flat_s[i], flat_a[i])
    m = _matrix_cache.get(key)
    if m is None:
      m = transform_matrix(flat_s[i], flat_a[i])
      if len(_matrix_cache) < 65536:
        _matrix_cache[key] = m
    flat_o[i] = m
  return out


def empty_batch(n_scenes, n_slots):
  b = {f: np.zeros((n_scenes, n_slots), np.float64) for f in F64_FIELDS}
  b['member'] = np.zeros((n_scenes, n_slots), np.uint32)
  b['shape'] = np.zeros((n_scenes, n_slots), np.uint8)
  b['pos_f32'] = np.zeros((n_scenes, n_slots), np.uint8)
  b['rgb'] = np.zeros((n_scenes, n_slots, 3), np.uint8)
  b['factors'] = np.zeros((n_scenes, n_slots, 5), np.float32)
  return b


def batch_from_factor_arrays(x, y, pos_f32, shape, angle, scale, c0, c1, c2, vx, vy, member, rgb):
  """All arguments are arrays shaped (n, S) (rgb: (n, S, 3)); shape 0 marks empty slots."""
  x = np.asarray(x, np.float64)
  b = empty_batch(*x.shape)
  b['x'][:], b['y'][:] = x, np.asarray(y, np.float64)
  b['vx'][:], b['vy'][:] = vx, vy
  b['member'][:], b['shape'][:], b['pos_f32'][:] = member, shape, pos_f32
  b['rgb'][:] = rgb
  m = transform_matrices(scale, angle)
  occupied = b['shape'] > 0
  for i, f in enumerate(('m00', 'm01', 'm10', 'm11')):
    b[f][:] = np.where(occupied, m[..., i], 0.0)
  b['factors'][..., 0], b['factors'][..., 1] = scale, angle
  b['factors'][..., 2], b['factors'][..., 3], b['factors'][..., 4] = c0, c1, c2
  return b
