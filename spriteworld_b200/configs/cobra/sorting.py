"""Sorting by hue: five hue bands, each with its own goal location; two sprites of
different bands per episode; the (red, blue) pair is held out for test
(reference: configs/cobra/sorting.py)."""
import itertools

import numpy as np

from spriteworld_b200 import factor_distributions as distribs
from spriteworld_b200 import sprite_generators as gen
from spriteworld_b200 import tasks
from spriteworld_b200.configs.cobra import common

MAX_EPISODE_LENGTH = 50
TERMINATE_DISTANCE = 0.075
RAW_REWARD_MULTIPLIER = 20.
NUM_TARGETS = 2

SUBTASKS = (
    {'distrib': distribs.Continuous('c0', 0.9, 1.), 'goal_position': np.array([0.75, 0.75])},    # red
    {'distrib': distribs.Continuous('c0', 0.55, 0.65), 'goal_position': np.array([0.75, 0.25])},  # blue
    {'distrib': distribs.Continuous('c0', 0.27, 0.37), 'goal_position': np.array([0.25, 0.75])},  # green
    {'distrib': distribs.Continuous('c0', 0.73, 0.83), 'goal_position': np.array([0.25, 0.25])},  # purple
    {'distrib': distribs.Continuous('c0', 0.1, 0.2), 'goal_position': np.array([0.5, 0.5])},      # yellow
)


def get_config(mode='train'):
  subtasks, one_sprite = [], []
  for sub in SUBTASKS:
    subtasks.append(tasks.FindGoalPosition(
        filter_distrib=sub['distrib'], goal_position=sub['goal_position'],
        terminate_distance=TERMINATE_DISTANCE, raw_reward_multiplier=RAW_REWARD_MULTIPLIER))
    factors = distribs.Product(tuple([sub['distrib']] + common.body_factors()))
    one_sprite.append(gen.generate_sprites(factors, num_sprites=1))
  combos = list(itertools.combinations(np.arange(len(SUBTASKS)), NUM_TARGETS))

  def pair(combo):
    return gen.chain_generators(*[one_sprite[i] for i in combo])

  if mode == 'train':
    sprite_gen = gen.sample_generator([pair(c) for c in combos[1:]])
  elif mode == 'test':
    sprite_gen = pair(combos[0])
  else:
    raise ValueError('Invalide mode {}.'.format(mode))
  sprite_gen = gen.shuffle(sprite_gen)
  task = tasks.MetaAggregated(subtasks, reward_aggregator='sum', termination_criterion='all')
  return common.config(task, sprite_gen, MAX_EPISODE_LENGTH, __file__, mode)
