"""Host-side view of one sprite.

In the engine a sprite is a row of the struct-of-arrays state on the GPU; this class is
the Python object `init_sprites()` callables hand to the Environment (same constructor
and properties as the reference's `spriteworld/sprite.py:45-214`) and what
`Environment.state()` returns.  The step path never touches these objects: they are
converted once per reset into scene arrays (scene.py) and uploaded.

Geometry is kept as a 2x2 float64 matrix applied to the unit-area vertex table of the
shape (what the reference caches as `_centered_path`, sprite.py:96-101).
"""
import collections
import math

import numpy as np

from spriteworld_b200 import constants

FACTOR_NAMES = ('x', 'y', 'shape', 'angle', 'scale', 'c0', 'c1', 'c2', 'x_vel', 'y_vel')

_MAX_TRIES = int(1e6)


def _rotation(degrees):
  th = math.radians(degrees)
  c, s = math.cos(th), math.sin(th)
  return np.array([[c, -s], [s, c]])


def _apply(matrix, pts):
  """Affine map as matplotlib evaluates it: (a*x + c*y) + e per coordinate."""
  x, y = pts[:, 0], pts[:, 1]
  out = np.empty_like(pts)
  out[:, 0] = (matrix[0, 0] * x + matrix[0, 1] * y) + 0.0
  out[:, 1] = (matrix[1, 0] * x + matrix[1, 1] * y) + 0.0
  return out


def polygon_contains(vertices, px, py):
  """Even-odd point-in-polygon on an implicitly closed float64 vertex loop
  (what matplotlib's Path.contains_point computes for sprite.py:113-115)."""
  v = np.asarray(vertices, dtype=np.float64)
  if len(v) < 3:
    return False
  w = np.roll(v, -1, axis=0)
  f0 = v[:, 1] >= py
  f1 = w[:, 1] >= py
  cross = ((w[:, 1] - py) * (v[:, 0] - w[:, 0]) >= (w[:, 0] - px) * (v[:, 1] - w[:, 1])) == f1
  return bool(np.count_nonzero((f0 != f1) & cross) & 1)


class Sprite(object):
  """A shape with position, pose, colour and velocity.  (x, y) are mathematical
  coordinates: (0, 0) is the lower-left corner of the frame."""

  def __init__(self, x=0.5, y=0.5, shape='square', angle=0, scale=0.1, c0=0, c1=0, c2=0,
               x_vel=0.0, y_vel=0.0):
    if shape not in constants.SHAPES:
      raise KeyError(shape)
    self._position = np.array([x, y])
    self._shape = shape
    self._angle = angle
    self._scale = scale
    self._color = (c0, c1, c2)
    self._velocity = (x_vel, y_vel)
    self._rebuild_path()

  # -- geometry -------------------------------------------------------------------
  def _rebuild_path(self):
    s = float(self._scale)
    rot = _rotation(self._angle)
    # rotate . scale: every entry is a single product
    self._matrix = np.array([[rot[0, 0] * s, rot[0, 1] * s], [rot[1, 0] * s, rot[1, 1] * s]])
    self._centred = _apply(self._matrix, constants.SHAPES[self._shape])

  @property
  def transform(self):
    """(m00, m01, m10, m11) of the map from the unit shape to sprite-centred coordinates,
    or None if a setter made the centred path no longer a linear image of the table."""
    return tuple(float(v) for v in self._matrix.reshape(-1))

  @property
  def centred_vertices(self):
    return self._centred

  @property
  def vertices(self):
    """World-coordinate vertices (sprite.py:128-133)."""
    out = np.empty_like(self._centred)
    out[:, 0] = self._centred[:, 0] + float(self._position[0])
    out[:, 1] = self._centred[:, 1] + float(self._position[1])
    return out

  def contains_point(self, point):
    d = np.asarray(point) - self._position
    return polygon_contains(self._centred, float(d[0]), float(d[1]))

  def sample_contained_position(self):
    lo = np.min(self._centred, axis=0)
    hi = np.max(self._centred, axis=0)
    for _ in range(_MAX_TRIES):
      candidate = self._position + np.random.uniform(lo, hi)
      if self.contains_point(candidate):
        return candidate
    raise ValueError('could not sample a point inside the sprite')

  # -- motion ---------------------------------------------------------------------
  def move(self, motion, keep_in_frame=False):
    self._position += motion
    if keep_in_frame:
      self._position = np.clip(self._position, 0.0, 1.0)

  def update_position(self, keep_in_frame=False):
    self.move(self._velocity, keep_in_frame=keep_in_frame)

  @property
  def out_of_frame(self):
    p = self._position
    return not bool(np.all(p >= [0., 0.]) and np.all(p <= [1., 1.]))

  # -- factors ----------------------------------------------------------------------
  x = property(lambda self: self._position[0])
  y = property(lambda self: self._position[1])
  c0 = property(lambda self: self._color[0])
  c1 = property(lambda self: self._color[1])
  c2 = property(lambda self: self._color[2])
  x_vel = property(lambda self: self._velocity[0])
  y_vel = property(lambda self: self._velocity[1])
  color = property(lambda self: self._color)
  position = property(lambda self: self._position)
  velocity = property(lambda self: self._velocity)

  @property
  def shape(self):
    return self._shape

  @shape.setter
  def shape(self, name):
    self._shape = name
    self._rebuild_path()

  @property
  def angle(self):
    return self._angle

  @angle.setter
  def angle(self, degrees):
    # incremental rotation of the cached path, like sprite.py:158-162
    rot = _rotation(degrees - self._angle)
    self._matrix = rot @ self._matrix
    self._centred = _apply(rot, self._centred)
    self._angle = degrees

  @property
  def scale(self):
    return self._scale

  @scale.setter
  def scale(self, value):
    # The reference rescales the cached path by (new - old), not new/old
    # (sprite.py:171-175; its tests encode this).  Kept for drop-in behaviour.
    k = value - self._scale
    self._matrix = self._matrix * k
    self._centred = self._centred * k
    self._scale = value

  @property
  def factors(self):
    return collections.OrderedDict((name, getattr(self, name)) for name in FACTOR_NAMES)

  def __repr__(self):
    return 'Sprite(%s)' % ', '.join('%s=%r' % kv for kv in self.factors.items())
