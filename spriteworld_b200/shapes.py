"""Unit-area vertex tables of the Spriteworld shapes.

The engine's geometry must equal the reference's bit for bit, so the tables are produced
with the same NumPy scalar expressions, in the same order, as the reference
(spriteworld/shapes.py:30-116): angle i*theta + theta_0, radius * (cos, sin), everything
divided by sqrt(area).  The result is checked against the reference's own tables stored
in the golden fixtures (tests/test_host_api.py).
"""
import numpy as np


def _ray(radius, angle):
  """radius * (cos angle, sin angle), evaluated like shapes.py:30-31."""
  return radius * np.array([np.cos(angle), np.sin(angle)])


def polygon(num_sides, theta_0=0.):
  """Regular polygon with `num_sides` vertices and area 1 (shapes.py:34-49)."""
  step = 2 * np.pi / num_sides
  rows = [_ray(1, k * step + theta_0) for k in range(num_sides)]
  area = num_sides * np.sin(step / 2) * np.cos(step / 2)
  return np.array(rows) / np.sqrt(area)


def star(num_sides, point_height=1, theta_0=0.):
  """Star with `num_sides` points and area 1 (shapes.py:52-74).

  Even vertices lie on the unit circle, odd vertices (the tips) at 1 + point_height.
  """
  tip_radius = 1 + point_height
  step = 2 * np.pi / num_sides
  out = np.empty([2 * num_sides, 2])
  for k in range(num_sides):
    out[2 * k] = _ray(1, k * step + theta_0)
    out[2 * k + 1] = _ray(tip_radius, (k + 0.5) * step + theta_0)
  area = tip_radius * num_sides * np.sin(step / 2)
  return np.array(out) / np.sqrt(area)


def spokes(num_sides, spoke_height=1, theta_0=0.):
  """Gear-like shape with rectangular spokes and area 1 (shapes.py:77-116)."""
  step = 2 * np.pi / num_sides
  out = np.empty([3 * num_sides, 2])
  arm = _ray(spoke_height, -0.5 * step + theta_0)
  for k in range(num_sides):
    hub = _ray(1, k * step + theta_0)
    out[3 * k] = arm + hub
    out[3 * k + 1] = hub
    arm = _ray(spoke_height, (k + 0.5) * step + theta_0)
    out[3 * k + 2] = arm + hub
  area = num_sides * np.sin(step / 2) * (2 + np.cos(step / 2))
  return np.array(out) / np.sqrt(area)
