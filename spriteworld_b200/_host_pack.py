"""ctypes binding of csrc/swb_host_pack.c: one native pass that packs a table of sampled sprites
into the scene arrays (the common case of scene.arrays_from_layout).  Host-side only -- no CUDA;
if the library has not been built, or the table is not of the plain kind it handles, pack()
returns False and the NumPy path runs (it is the specification; tests compare the two)."""
import ctypes
import os

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc', 'libswb_host.so')
_lib = None
_tried = False

_CODES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int64): 2}
_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64


class _Args(ctypes.Structure):
  _fields_ = (
      [(n, _vp) for n in ('x', 'y', 'scale', 'angle', 'c0', 'c1', 'c2', 'vx', 'vy')] +
      [(n, _i32) for n in ('x_code', 'y_code', 'scale_code', 'angle_code', 'c_code', 'vx_code',
                           'vy_code', 'pos_f32', 'rgb_mode')] +
      [('transform', _vp), ('transform_stride', _i64), ('member', _vp), ('shape_ids', _vp), ('shape_ptrs', _vp),
       ('shape_key', _vp), ('shape_val', _vp), ('n_shape_keys', _i64)] +
      [(n, _vp) for n in ('o_x', 'o_y', 'o_m00', 'o_m01', 'o_m10', 'o_m11', 'o_vx', 'o_vy',
                          'o_member', 'o_shape', 'o_pos_f32', 'o_rgb', 'o_factors')])


def lib_path():
  return _LIB_PATH


def load():
  """The library, or None if it has not been built (python -m spriteworld_b200.build)."""
  global _lib, _tried
  if _tried:
    return _lib
  _tried = True
  if os.environ.get('SPRITEWORLD_B200_NO_HOST_PACK') or not os.path.exists(_LIB_PATH):
    return None
  try:
    L = ctypes.CDLL(_LIB_PATH)
    L.swb_pack_scenes.argtypes = [ctypes.POINTER(_Args), _i64, _vp, _vp]
    L.swb_pack_scenes.restype = ctypes.c_int
    L.swb_distinct_pointers.argtypes = [_vp, _i64, _i64, _vp, _vp]
    L.swb_distinct_pointers.restype = _i64
    if L.swb_host_pack_version() != 1:
      return None
    _lib = L
  except OSError:
    _lib = None
  return _lib


def _ptr(a):
  return a.ctypes.data


def pack(table, filters, color_to_rgb, rows, dst, out):
  """Fills `out` (scene.empty_batch arrays) from table rows `rows` at flat slots `dst` (None:
  slot t for item t).  Returns False if this table is not handled here (nothing written)."""
  L = load()
  if L is None:
    return False
  from spriteworld_b200 import constants, scene
  from spriteworld_b200.renderers import color_maps
  if '_transform' in table.columns or '_pos_f32' in table.columns:
    return False
  cols = scene._full_columns(table)
  keep = []   # arrays the C call reads must stay alive until it returns

  def numeric(name):
    c = cols[name]
    if c.dtype not in _CODES or not c.flags.c_contiguous:
      return None
    keep.append(c)
    return c

  x, y, sc, an = numeric('x'), numeric('y'), numeric('scale'), numeric('angle')
  c0, c1, c2 = numeric('c0'), numeric('c1'), numeric('c2')
  vx, vy = numeric('x_vel'), numeric('y_vel')
  if any(c is None for c in (x, y, sc, an, c0, c1, c2, vx, vy)):
    return False
  if not (c0.dtype == c1.dtype == c2.dtype):
    return False
  if color_to_rgb is None:
    rgb_mode = 0
  elif color_to_rgb is color_maps.hsv_to_rgb and c0.dtype in (np.dtype(np.float32), np.dtype(np.float64)):
    rgb_mode = 1
  else:
    return False
  a = _Args()
  shape = cols['shape']
  if shape.dtype == object:
    if not shape.flags.c_contiguous or len(shape) == 0:
      return False
    keys = np.empty(32, np.uintp)
    first = np.empty(32, np.int64)
    k = L.swb_distinct_pointers(_ptr(shape), len(shape), 32, _ptr(keys), _ptr(first))
    if k < 0:
      return False
    try:
      vals = np.array([int(constants.ShapeType[str(shape[int(i)])]) for i in first[:k]], np.uint8)
    except KeyError:
      return False
    keep += [shape, keys, vals]
    a.shape_ids, a.shape_ptrs, a.shape_key, a.shape_val, a.n_shape_keys = (
        None, _ptr(shape), _ptr(keys), _ptr(vals), int(k))
  elif shape.dtype == np.dtype(np.int64) and shape.flags.c_contiguous:
    keep.append(shape)
    a.shape_ids, a.shape_ptrs, a.shape_key, a.shape_val, a.n_shape_keys = _ptr(shape), None, None, None, 0
  else:
    return False
  # the parts that stay in NumPy: task filters (arbitrary distributions) and the transforms
  # (libm through math.cos / math.sin, once per distinct (scale, angle))
  member = np.zeros(table.rows, np.uint32)
  for bit, f in enumerate(filters):
    member |= np.asarray(f.contains_batch(cols), bool).astype(np.uint32) << np.uint32(bit)
  sb = sc.view(np.int32 if sc.dtype == np.float32 else np.int64)
  ab = an.view(np.int32 if an.dtype == np.float32 else np.int64)
  if sb.min() == sb.max() and ab.min() == ab.max():    # one (scale, angle) bit pattern for all
    transform = np.array(scene.transform_matrix(float(sc[0]), float(an[0])), np.float64)
    a.transform_stride = 0
  else:
    transform = np.ascontiguousarray(
        scene.transform_matrices(scene._numeric(sc), scene._numeric(an)), dtype=np.float64)
    a.transform_stride = 4
  rows = np.ascontiguousarray(rows, dtype=np.int64)
  dst_arr = None if dst is None else np.ascontiguousarray(dst, dtype=np.int64)
  keep += [member, transform, rows, dst_arr]
  a.x, a.y, a.scale, a.angle = _ptr(x), _ptr(y), _ptr(sc), _ptr(an)
  a.c0, a.c1, a.c2, a.vx, a.vy = _ptr(c0), _ptr(c1), _ptr(c2), _ptr(vx), _ptr(vy)
  a.x_code, a.y_code, a.scale_code, a.angle_code = (_CODES[c.dtype] for c in (x, y, sc, an))
  a.c_code, a.vx_code, a.vy_code = _CODES[c0.dtype], _CODES[vx.dtype], _CODES[vy.dtype]
  a.pos_f32 = int(x.dtype == np.float32 and y.dtype == np.float32)
  a.rgb_mode = rgb_mode
  a.transform, a.member = _ptr(transform), _ptr(member)
  for name in ('x', 'y', 'm00', 'm01', 'm10', 'm11', 'vx', 'vy', 'member', 'shape', 'pos_f32', 'rgb',
               'factors'):
    arr = out[name]
    if not arr.flags.c_contiguous:
      return False
    setattr(a, 'o_' + name, _ptr(arr))
  rc = L.swb_pack_scenes(ctypes.byref(a), len(rows), _ptr(rows), None if dst_arr is None else _ptr(dst_arr))
  return rc == 0
