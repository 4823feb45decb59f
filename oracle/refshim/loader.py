"""Import google-deepmind/spriteworld from /root/reference without editing it.

The reference does not import in this image: `matplotlib`,
`dm_env`, `gym`, `mock` are absent and NumPy 2.x / Pillow 12 removed four
aliases it uses.  This loader supplies

  * alias patches: np.cast, np.asscalar, np.object, PIL.Image.ANTIALIAS,
    sys.modules['mock'];
  * stand-in packages `dm_env` and `matplotlib.{path,transforms}` (only when the
    real ones are absent) from oracle/refshim/standins/.

The reference's own source files are imported as they lie.
"""
import os
import sys
import unittest.mock

REFERENCE_ROOT = os.environ.get('SPRITEWORLD_REFERENCE', '/root/reference')
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'standins')


def reference_available():
  return os.path.isdir(os.path.join(REFERENCE_ROOT, 'spriteworld'))


class _Cast(dict):
  """np.cast[dtype](x) as removed in NumPy 2.0 (factor_distributions.py:102)."""

  def __missing__(self, dtype):
    import numpy as np
    fn = lambda x, _d=np.dtype(dtype): np.asarray(x).astype(_d)[()]
    self[dtype] = fn
    return fn


def _patch_aliases():
  import numpy as np
  from PIL import Image
  if not hasattr(np, 'cast'):
    np.cast = _Cast()
  if not hasattr(np, 'asscalar'):
    np.asscalar = lambda a: a.item()
  if 'object' not in np.__dict__:
    np.object = object
  if not hasattr(Image, 'ANTIALIAS'):
    Image.ANTIALIAS = Image.LANCZOS  # ANTIALIAS was the old name of LANCZOS
  sys.modules.setdefault('mock', unittest.mock)


def _have(mod):
  import importlib.util
  try:
    return importlib.util.find_spec(mod) is not None
  except (ImportError, ValueError):
    return False


def load_reference():
  """Returns the imported reference `spriteworld` package."""
  if not reference_available():
    raise ImportError('reference not found at %s' % REFERENCE_ROOT)
  _patch_aliases()
  if not (_have('matplotlib') and _have('dm_env')):
    if _STANDINS not in sys.path:
      sys.path.append(_STANDINS)  # appended: real packages win if present
  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
  import spriteworld  # noqa: F401  (the reference package)
  from spriteworld import (action_spaces, environment, factor_distributions,  # noqa
                           sprite, sprite_generators, tasks, renderers)
  return spriteworld
