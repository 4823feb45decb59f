"""Rates of BatchedEnvironment.step for several refill configurations (development aid).

  python tools/api_probe.py                # C2 through the plugin API, in-process sampler
  python tools/api_probe.py --procs 8      # with worker processes (_sampler_pool)
  python tools/api_probe.py --pool-only    # the sampler pool alone, no GPU

Prints env-steps/s, the host time per scene and how long steps had to wait for scenes.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pool_only():
  from concurrent.futures import ThreadPoolExecutor
  from spriteworld_b200 import _sampler_pool, scene, sprite_generators, workloads
  from spriteworld_b200.renderers import color_maps
  cfg = workloads.WORKLOADS['c2']().plugin_config()
  _, filters = cfg['task'].compile()
  n = 11700
  t0 = time.perf_counter()
  for r in range(3):
    lay = sprite_generators.batch_of(cfg['init_sprites'], n, np.random.RandomState(r))
    scene.arrays_from_layout(lay, 5, filters, color_maps.hsv_to_rgb)
  print('in-process: %.2f us/scene' % ((time.perf_counter() - t0) / 3 / n * 1e6), flush=True)
  for procs in (2, 4, 8):
    pool = _sampler_pool.SamplerPool(procs, cfg['init_sprites'], 5, filters, color_maps.hsv_to_rgb)
    for i in range(procs):
      pool.sample(i, 10, 1)
    with ThreadPoolExecutor(procs) as ex:
      t0 = time.perf_counter()
      for r in range(5):
        list(ex.map(lambda i: pool.sample(i, n // procs, r * 7 + i), range(procs)))
      dt = (time.perf_counter() - t0) / 5
    print('%d procs: %.2f ms per %d scenes = %.2f us/scene' % (procs, dt * 1e3, n, dt / n * 1e6), flush=True)
    pool.close()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--pool-only', action='store_true')
  ap.add_argument('--procs', type=int, default=0)
  ap.add_argument('--threads', type=int, default=1)
  ap.add_argument('--pool-depth', type=int, default=32)
  ap.add_argument('--steps', type=int, default=300)
  ap.add_argument('--repeat', type=int, default=2)
  args = ap.parse_args()
  if args.pool_only:
    return pool_only()
  import torch
  from spriteworld_b200 import environment, workloads
  wl = workloads.WORKLOADS['c2']()
  E = wl.n_envs
  for _ in range(args.repeat):
    env = environment.BatchedEnvironment(
        n_envs=E, device=0, rng=np.random.RandomState(1), pool_depth=args.pool_depth,
        refill_threads=args.threads, refill_procs=args.procs, **wl.plugin_config())
    acts = torch.from_numpy(wl.sample_actions(np.random.RandomState(5), 16, E)).to(env.engine.device)
    for i in range(60):
      env.step(acts[i % 16])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
      env.step(acts[i % 16])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = env.refill_stats()
    print('procs %d threads %d K %d: %.2f M env-steps/s, %.3f ms/step; refills %d scenes %d host %.3fs '
          '(%.2f us/scene) blocked %d %.3fs'
          % (args.procs, args.threads, args.pool_depth, E * args.steps / dt / 1e6, 1e3 * dt / args.steps,
             st['refills'], st['scenes'], st['host_seconds'],
             1e6 * st['host_seconds'] / max(st['scenes'], 1), st['blocked'], st['blocked_seconds']), flush=True)
    env.close()


if __name__ == '__main__':
  main()
