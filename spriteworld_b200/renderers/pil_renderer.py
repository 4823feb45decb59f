"""PILRenderer: anti-aliased polygon rendering of sprites to uint8 RGB.

Same constructor as the reference's `renderers/pil_renderer.py:32-65`.  The class only
*describes* the renderer (size, anti-aliasing factor, background, colour map); frames are
produced by the render kernel (csrc/swb_render.cuh), which reproduces what the
reference gets from Pillow -- `ImageDraw.polygon` on an anti_aliasing-times larger canvas
and `Image.resize(..., ANTIALIAS)` -- bit for bit, without materialising the canvas.
`render(sprites)` on a Python sprite list renders through a one-env engine on the GPU.
"""
import numpy as np

from spriteworld_b200._dm_env import specs
from spriteworld_b200.renderers import abstract_renderer


class PILRenderer(abstract_renderer.AbstractRenderer):

  def __init__(self, image_size=(64, 64), anti_aliasing=1, bg_color=None, color_to_rgb=None):
    """Args as in the reference.  Note image_size is handed to Pillow as (width, height)
    there (pil_renderer.py:50-51,59,84), so frames are (image_size[1], image_size[0], 3)."""
    self._image_size = tuple(int(v) for v in image_size)
    self._anti_aliasing = int(anti_aliasing)
    self._canvas_size = (self._anti_aliasing * self._image_size[0],
                         self._anti_aliasing * self._image_size[1])
    self._color_to_rgb = color_to_rgb           # None: colours already are RGB
    self._bg_color = (0, 0, 0) if bg_color is None else tuple(int(c) for c in bg_color)
    self._observation_spec = specs.Array(shape=self._image_size + (3,), dtype=np.uint8)

  # -- description consumed by the engine ------------------------------------------
  @property
  def width(self):
    return self._image_size[0]

  @property
  def height(self):
    return self._image_size[1]

  @property
  def anti_aliasing(self):
    return self._anti_aliasing

  @property
  def bg_color(self):
    return self._bg_color

  @property
  def color_to_rgb(self):
    return self._color_to_rgb

  # -- plugin protocol ---------------------------------------------------------------
  def render(self, sprites=(), global_state=None):
    from spriteworld_b200 import _direct
    return _direct.render(self, list(sprites))

  def observation_spec(self):
    return self._observation_spec
