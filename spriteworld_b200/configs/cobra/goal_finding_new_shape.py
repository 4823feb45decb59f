"""Goal finding, generalisation to new shapes: one sprite, squares in train, triangles and
circles in test; every sprite must reach the goal (no filter)
(reference: configs/cobra/goal_finding_new_shape.py)."""
from spriteworld_b200 import factor_distributions as distribs
from spriteworld_b200 import sprite_generators as gen
from spriteworld_b200 import tasks
from spriteworld_b200.configs.cobra import common

TERMINATE_DISTANCE = 0.075
NUM_TARGETS = 1
MODES_SHAPES = {
    'train': distribs.Discrete('shape', ['square']),
    'test': distribs.Discrete('shape', ['triangle', 'circle']),
}


def get_config(mode='train'):
  factors = distribs.Product([
      MODES_SHAPES[mode],
      distribs.Continuous('x', 0.1, 0.9),
      distribs.Continuous('y', 0.1, 0.9),
      distribs.Discrete('scale', [0.13]),
      distribs.Continuous('c0', 0., 0.4),
      distribs.Continuous('c1', 0.3, 1.),
      distribs.Continuous('c2', 0.9, 1.),
  ])
  sprite_gen = gen.shuffle(gen.generate_sprites(factors, num_sprites=NUM_TARGETS))
  task = tasks.FindGoalPosition(terminate_distance=TERMINATE_DISTANCE)
  return common.config(task, sprite_gen, 20, __file__, mode)
