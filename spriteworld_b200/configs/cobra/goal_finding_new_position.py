"""Goal finding, generalisation to new positions: in train the target never starts in the
upper-right quadrant [0.5, 0.9) x [0.5, 0.9), in test it always does
(reference: configs/cobra/goal_finding_new_position.py)."""
from spriteworld_b200 import factor_distributions as distribs
from spriteworld_b200 import sprite_generators as gen
from spriteworld_b200 import tasks
from spriteworld_b200.configs.cobra import common

TERMINATE_DISTANCE = 0.075
NUM_TARGETS = 1
NUM_DISTRACTORS = 1


def _box(lo, hi):
  return distribs.Product((distribs.Continuous('x', lo, hi), distribs.Continuous('y', lo, hi)))


MODES_TARGET_POSITIONS = {
    'train': distribs.SetMinus(_box(0.1, 0.9), _box(0.5, 0.9)),
    'test': _box(0.5, 0.9),
}


def get_config(mode='train'):
  appearance = distribs.Product(common.body_factors()[2:])   # shape, scale, c1, c2
  target_hue = distribs.Continuous('c0', 0., 0.4)
  distractor_hue = distribs.Continuous('c0', 0.5, 0.9)
  target = distribs.Product([MODES_TARGET_POSITIONS[mode], target_hue, appearance])
  distractor = distribs.Product([distribs.Continuous('x', 0.1, 0.9),
                                 distribs.Continuous('y', 0.1, 0.9), distractor_hue, appearance])
  sprite_gen = gen.shuffle(gen.chain_generators(
      gen.generate_sprites(target, num_sprites=NUM_TARGETS),
      gen.generate_sprites(distractor, num_sprites=NUM_DISTRACTORS)))
  task = tasks.FindGoalPosition(filter_distrib=target_hue, terminate_distance=TERMINATE_DISTANCE)
  return common.config(task, sprite_gen, 20, __file__, mode)
