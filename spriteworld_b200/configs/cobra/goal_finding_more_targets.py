"""Goal finding, generalisation to more targets: 1 target in train, 2 in test, always 2
distractors (reference: configs/cobra/goal_finding_more_targets.py)."""
from spriteworld_b200 import factor_distributions as distribs
from spriteworld_b200 import sprite_generators as gen
from spriteworld_b200 import tasks
from spriteworld_b200.configs.cobra import common

TERMINATE_DISTANCE = 0.075
NUM_DISTRACTORS = 2
MODES_NUM_TARGETS = {'train': 1, 'test': 2}


def get_config(mode='train'):
  shared = distribs.Product(common.body_factors())
  target_hue = distribs.Continuous('c0', 0., 0.4)
  distractor_hue = distribs.Continuous('c0', 0.5, 0.9)
  targets = gen.generate_sprites(distribs.Product([target_hue, shared]),
                                 num_sprites=MODES_NUM_TARGETS[mode])
  distractors = gen.generate_sprites(distribs.Product([distractor_hue, shared]),
                                     num_sprites=NUM_DISTRACTORS)
  sprite_gen = gen.shuffle(gen.chain_generators(targets, distractors))
  task = tasks.FindGoalPosition(filter_distrib=target_hue, terminate_distance=TERMINATE_DISTANCE)
  return common.config(task, sprite_gen, 20, __file__, mode)
