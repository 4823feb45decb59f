"""Pieces shared by every COBRA task (reference: configs/cobra/common.py:26-38)."""
from spriteworld_b200 import action_spaces
from spriteworld_b200 import factor_distributions as distribs
from spriteworld_b200 import renderers as sw_renderers

SHAPES = ('square', 'triangle', 'circle')


def action_space():
  return action_spaces.SelectMove(scale=0.25)


def renderers():
  return {
      'image': sw_renderers.PILRenderer(
          image_size=(64, 64), anti_aliasing=5,
          color_to_rgb=sw_renderers.color_maps.hsv_to_rgb),
  }


def body_factors(shapes=SHAPES, position=(0.1, 0.9), scale=0.13):
  """Position, shape, scale, saturation and value factors every COBRA sprite shares."""
  lo, hi = position
  return [
      distribs.Continuous('x', lo, hi),
      distribs.Continuous('y', lo, hi),
      distribs.Discrete('shape', list(shapes)),
      distribs.Discrete('scale', [scale]),
      distribs.Continuous('c1', 0.3, 1.),
      distribs.Continuous('c2', 0.9, 1.),
  ]


def config(task, sprite_gen, max_episode_length, source_file, mode):
  import os
  return {
      'task': task,
      'action_space': action_space(),
      'renderers': renderers(),
      'init_sprites': sprite_gen,
      'max_episode_length': max_episode_length,
      'metadata': ({'name': os.path.basename(source_file), 'mode': mode} if mode is not None
                   else {'name': os.path.basename(source_file)}),
  }
