"""Clustering by hue: two sprites from each of two hue bands; reward is the inverse
Davies-Bouldin index (reference: configs/cobra/clustering.py)."""
from spriteworld_b200 import factor_distributions as distribs
from spriteworld_b200 import sprite_generators as gen
from spriteworld_b200 import tasks
from spriteworld_b200.configs.cobra import common

NUM_SPRITES_PER_CLUSTER = 2
MAX_EPISODE_LENGTH = 50

CLUSTERS_DISTS = {
    'red': distribs.Continuous('c0', 0.9, 1.),
    'blue': distribs.Continuous('c0', 0.55, 0.65),
    'green': distribs.Continuous('c0', 0.27, 0.37),
    'yellow': distribs.Continuous('c0', 0.1, 0.2),
}
MODES = {'train': ('blue', 'green'), 'test': ('red', 'yellow')}


def get_config(mode='train'):
  hues = [CLUSTERS_DISTS[name] for name in MODES[mode]]
  other = distribs.Product(common.body_factors())
  per_cluster = [gen.generate_sprites(distribs.Product((other, hue)),
                                      num_sprites=NUM_SPRITES_PER_CLUSTER) for hue in hues]
  sprite_gen = gen.shuffle(gen.chain_generators(*per_cluster))
  task = tasks.Clustering(hues, terminate_bonus=0., reward_range=10.)
  return common.config(task, sprite_gen, MAX_EPISODE_LENGTH, __file__, mode)
