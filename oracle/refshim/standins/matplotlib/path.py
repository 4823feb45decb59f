"""Stand-in for matplotlib.path.Path: vertices + even-odd contains_point.

Follows matplotlib's C++ `point_in_path_impl` (src/_path.h): float64, the vertex
loop is implicitly closed, an edge toggles the inside flag when its endpoints
straddle the horizontal through the point and the +X ray crosses it.
"""
import numpy as np


class Path(object):

  def __init__(self, vertices, codes=None, _interpolation_steps=1):
    self.vertices = np.asarray(vertices, dtype=np.float64)
    self.codes = codes

  def contains_point(self, point, transform=None, radius=0.0):
    tx = float(point[0])
    ty = float(point[1])
    v = self.vertices
    n = len(v)
    if n < 3:
      return False
    inside = False
    sx, sy = float(v[0][0]), float(v[0][1])
    vx0, vy0 = sx, sy
    yflag0 = vy0 >= ty
    for i in range(1, n + 1):
      if i < n:
        vx1, vy1 = float(v[i][0]), float(v[i][1])
      else:
        vx1, vy1 = sx, sy  # closing edge
      yflag1 = vy1 >= ty
      if yflag0 != yflag1:
        if ((vy1 - ty) * (vx0 - vx1) >= (vx1 - tx) * (vy0 - vy1)) == yflag1:
          inside = not inside
      yflag0 = yflag1
      vx0, vy0 = vx1, vy1
    return inside
