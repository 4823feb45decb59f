"""Colour-space conversions for rendering (reference: renderers/color_maps.py:26-28).

`hsv_to_rgb(c)` is the per-sprite callable a PILRenderer takes as `color_to_rgb`;
`hsv_to_rgb_batch` is its vectorised form used when scenes are sampled in bulk.  Both
reproduce `colorsys.hsv_to_rgb` followed by `(255 * rgb).astype(uint8)` (truncation),
evaluated in float32 when the colour factors are NumPy float32 scalars (as sampled by
`Continuous(..., dtype='float32')`) and in float64 otherwise.
"""
import colorsys

import numpy as np


def hsv_to_rgb(c):
  """HSV tuple -> uint8 RGB tuple."""
  return tuple((255 * np.array(colorsys.hsv_to_rgb(*c))).astype(np.uint8))


# which of (v, q, p, t) each channel takes in colorsys.hsv_to_rgb's cases i = 0..5
_R_CASE = np.array([0, 1, 2, 2, 3, 0])
_G_CASE = np.array([3, 0, 0, 1, 2, 2])
_B_CASE = np.array([2, 2, 3, 0, 0, 1])


def _hsv_arrays(h, s, v):
  """colorsys.hsv_to_rgb on arrays of one float dtype (elementwise IEEE ops, no fusion)."""
  dt = h.dtype.type
  one, six = dt(1.0), dt(6.0)
  h6 = h * six
  i = h6.astype(np.int64)  # int() truncates
  f = h6 - i.astype(h.dtype)
  p = v * (one - s)
  q = v * (one - s * f)
  t = v * (one - s * (one - f))
  i = i % 6
  # colorsys' six cases as one gather per channel from (v, q, p, t) laid end to end
  # (np.choose over six arrays is four times slower)
  flat = i.reshape(-1)
  n = flat.shape[0]
  vals = np.concatenate([v.reshape(-1), q.reshape(-1), p.reshape(-1), t.reshape(-1)])
  base = np.arange(n)
  r = vals[_R_CASE[flat] * n + base].reshape(i.shape)
  g = vals[_G_CASE[flat] * n + base].reshape(i.shape)
  b = vals[_B_CASE[flat] * n + base].reshape(i.shape)
  grey = s == 0
  if grey.any():
    r, g, b = np.where(grey, v, r), np.where(grey, v, g), np.where(grey, v, b)
  return r, g, b


def hsv_to_rgb_batch(c0, c1, c2, is_f32):
  """Vectorised hsv_to_rgb.

  Args:
    c0, c1, c2: float64 arrays (values of the colour factors).
    is_f32: bool array, True where all three factors of the sprite were float32 scalars.
  Returns:
    uint8 array shaped c0.shape + (3,).
  """
  c0, c1, c2 = (np.asarray(a, np.float64) for a in (c0, c1, c2))
  is_f32 = np.broadcast_to(np.asarray(is_f32, bool), c0.shape)
  out = np.zeros(c0.shape + (3,), np.uint8)
  uniform = None if is_f32.size == 0 else (True if is_f32.all() else (False if not is_f32.any() else None))
  if uniform is not None:   # one float type for every sprite (the usual case): no masks
    dt = np.float32 if uniform else np.float64
    r, g, b = _hsv_arrays(c0.astype(dt), c1.astype(dt), c2.astype(dt))
    scale = dt(255)
    out[..., 0] = (scale * r).astype(np.uint8)
    out[..., 1] = (scale * g).astype(np.uint8)
    out[..., 2] = (scale * b).astype(np.uint8)
    return out
  for f32 in (True, False):
    sel = is_f32 == f32
    if not sel.any():
      continue
    dt = np.float32 if f32 else np.float64
    r, g, b = _hsv_arrays(c0[sel].astype(dt), c1[sel].astype(dt), c2[sel].astype(dt))
    scale = dt(255)
    out[sel, 0] = (scale * r).astype(np.uint8)
    out[sel, 1] = (scale * g).astype(np.uint8)
    out[sel, 2] = (scale * b).astype(np.uint8)
  return out
