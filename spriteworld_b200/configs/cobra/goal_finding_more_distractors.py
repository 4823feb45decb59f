"""Goal finding, generalisation to more distractors: 2 targets, 1 distractor in train and 2
in test (reference: configs/cobra/goal_finding_more_distractors.py)."""
from spriteworld_b200 import factor_distributions as distribs
from spriteworld_b200 import sprite_generators as gen
from spriteworld_b200 import tasks
from spriteworld_b200.configs.cobra import common

TERMINATE_DISTANCE = 0.075
NUM_TARGETS = 2
MODES_NUM_DISTRACTORS = {'train': 1, 'test': 2}


def get_config(mode='train'):
  shared = distribs.Product(common.body_factors())
  target_hue = distribs.Continuous('c0', 0., 0.4)
  distractor_hue = distribs.Continuous('c0', 0.5, 0.9)
  targets = gen.generate_sprites(distribs.Product([target_hue, shared]), num_sprites=NUM_TARGETS)
  distractors = gen.generate_sprites(distribs.Product([distractor_hue, shared]),
                                     num_sprites=MODES_NUM_DISTRACTORS[mode])
  sprite_gen = gen.shuffle(gen.chain_generators(targets, distractors))
  task = tasks.FindGoalPosition(filter_distrib=target_hue, terminate_distance=TERMINATE_DISTANCE)
  return common.config(task, sprite_gen, 20, __file__, mode)
