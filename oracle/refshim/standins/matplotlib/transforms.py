"""Stand-in for matplotlib.transforms.Affine2D (scale, rotate_deg, translate, +).

Matrix composition and `transform_path` follow matplotlib: `A + B` applies A
then B (matrix B.A); a point maps as  x' = (a*x + c*y) + e,  y' = (b*x + d*y) + f
(two products, then two adds, float64, no FMA) as in `affine_transform_2d`.
"""
import math
import numpy as np
from matplotlib.path import Path


class Affine2D(object):

  def __init__(self, matrix=None):
    self._mtx = np.identity(3) if matrix is None else np.array(matrix, float)

  def get_matrix(self):
    return self._mtx

  def scale(self, sx, sy=None):
    if sy is None:
      sy = sx
    m = np.array([[sx, 0.0, 0.0], [0.0, sy, 0.0], [0.0, 0.0, 1.0]], float)
    self._mtx = np.dot(m, self._mtx)
    return self

  def rotate(self, theta):
    a = math.cos(theta)
    b = math.sin(theta)
    m = np.array([[a, -b, 0.0], [b, a, 0.0], [0.0, 0.0, 1.0]], float)
    self._mtx = np.dot(m, self._mtx)
    return self

  def rotate_deg(self, degrees):
    return self.rotate(math.radians(degrees))

  def translate(self, tx, ty):
    m = np.array([[1.0, 0.0, tx], [0.0, 1.0, ty], [0.0, 0.0, 1.0]], float)
    self._mtx = np.dot(m, self._mtx)
    return self

  def __add__(self, other):
    return Affine2D(np.dot(other.get_matrix(), self._mtx))

  def transform(self, points):
    pts = np.asarray(points, dtype=np.float64)
    m = self._mtx
    a, c, e = float(m[0, 0]), float(m[0, 1]), float(m[0, 2])
    b, d, f = float(m[1, 0]), float(m[1, 1]), float(m[1, 2])
    out = np.empty_like(pts)
    x = pts[:, 0]
    y = pts[:, 1]
    out[:, 0] = (a * x + c * y) + e
    out[:, 1] = (b * x + d * y) + f
    return out

  def transform_path(self, path):
    return Path(self.transform(path.vertices), path.codes)
