"""Times the UNMODIFIED reference's Environment.step (build container only).

  python tools/time_reference.py

Loads google-deepmind/spriteworld from /root/reference through oracle/refshim (alias patches
for NumPy 2 / Pillow 12; stand-ins for the absent matplotlib and dm_env, so the hit tests run
through a Python restatement of matplotlib's point_in_path -- slower than matplotlib's C) and
steps (a) configs/cobra/goal_finding.py (BASELINE C1: 1 env, 64x64 PILRenderer, aa=5) and (b)
the C2 scene mix (2 targets + 3 distractors, 5 sprites) with random SelectMove actions, one
process, one core.  Prints env-steps/s and the share of the step spent in PILRenderer.render.
The numbers go into BASELINE.md section 3 beside the oracle port's rate.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(env, steps, warm=50):
  rng = np.random.RandomState(0)
  env.reset()
  for _ in range(warm):
    env.step(rng.uniform(0, 1, 4))
  t0 = time.perf_counter()
  for _ in range(steps):
    env.step(rng.uniform(0, 1, 4))
  return steps / (time.perf_counter() - t0)


def main():
  from oracle.refshim import loader
  sw = loader.load_reference()
  from spriteworld import environment, renderers, sprite_generators, tasks
  from spriteworld import factor_distributions as distribs
  from spriteworld.configs.cobra import common, goal_finding_more_targets
  import importlib
  np.random.seed(0)
  out = {}
  # (a) the shipped config (BASELINE configs[0])
  cfg = importlib.import_module('spriteworld.configs.cobra.goal_finding_new_position').get_config('train')
  env = environment.Environment(**cfg)
  out['c1_goal_finding_new_position'] = run(env, 1500)
  # (b) C2's scene mix: 2 targets + 3 distractors (goal_finding_more_targets.py:54-86, 5 sprites)
  shared = [distribs.Continuous('x', 0.1, 0.9), distribs.Continuous('y', 0.1, 0.9),
            distribs.Discrete('shape', ['square', 'triangle', 'circle']),
            distribs.Discrete('scale', [0.13]), distribs.Continuous('c1', 0.3, 1.),
            distribs.Continuous('c2', 0.9, 1.)]
  th, dh = distribs.Continuous('c0', 0., 0.4), distribs.Continuous('c0', 0.5, 0.9)
  gen = sprite_generators.shuffle(sprite_generators.chain_generators(
      sprite_generators.generate_sprites(distribs.Product([th] + shared), num_sprites=2),
      sprite_generators.generate_sprites(distribs.Product([dh] + shared), num_sprites=3)))
  task = tasks.FindGoalPosition(filter_distrib=th, terminate_distance=0.075)
  cfg2 = dict(task=task, action_space=common.action_space(), renderers=common.renderers(),
              init_sprites=gen, max_episode_length=20)
  env2 = environment.Environment(**cfg2)
  out['c2_mix_5_sprites'] = run(env2, 1500)
  # share of PILRenderer.render in a step
  r = common.renderers()['image']
  sprites = gen()
  t0 = time.perf_counter()
  for _ in range(500):
    r.render(sprites)
  out['pil_render_ms_5_sprites'] = 1e3 * (time.perf_counter() - t0) / 500
  import PIL
  print('reference Environment.step, one core, Pillow %s, NumPy %s' % (PIL.__version__, np.__version__))
  for k, v in out.items():
    print('  %-32s %.1f' % (k, v))


if __name__ == '__main__':
  main()
