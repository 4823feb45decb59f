"""Rates of BatchedEnvironment.step for several refill configurations (development aid).

  python tools/api_probe.py [--pool-only]
"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pool_only():
  from concurrent.futures import ThreadPoolExecutor
  from spriteworld_b200 import workloads, _sampler_pool, scene, sprite_generators
  from spriteworld_b200.renderers import color_maps
  cfg = workloads.WORKLOADS['c2']().plugin_config()
  nodes, filters = cfg['task'].compile()
  N = 11700
  t0 = time.perf_counter()
  for r in range(3):
    lay = sprite_generators.batch_of(cfg['init_sprites'], N, np.random.RandomState(r))
    scene.arrays_from_layout(lay, 5, filters, color_maps.hsv_to_rgb)
  print('in-process: %.2f us/scene' % ((time.perf_counter() - t0) / 3 / N * 1e6), flush=True)
  for P in (2, 4, 8):
    pool = _sampler_pool.SamplerPool(P, cfg['init_sprites'], 5, filters, color_maps.hsv_to_rgb)
    for i in range(P):
      pool.sample(i, 10, 1)
    with ThreadPoolExecutor(P) as ex:
      t0 = time.perf_counter()
      for r in range(5):
        list(ex.map(lambda i: pool.sample(i, N // P, r * 7 + i), range(P)))
      dt = (time.perf_counter() - t0) / 5
    print('%d procs: %.2f ms per %d scenes = %.2f us/scene' % (P, dt * 1e3, N, dt / N * 1e6), flush=True)
    pool.close()


def main():
  if '--pool-only' in sys.argv:
    return pool_only()
  import torch
  if '--cuda-then-pool' in sys.argv:
    torch.zeros(1, device='cuda')
    print('affinity after CUDA init: %d cpus' % len(os.sched_getaffinity(0)), flush=True)
    return pool_only()
  def cpu_stat():
    try:
      return dict(l.split() for l in open('/sys/fs/cgroup/cpu.stat').read().strip().splitlines())
    except OSError:
      return {}
  print('torch threads', torch.get_num_threads(), 'interop', torch.get_num_interop_threads(), flush=True)
  if '--one-thread' in sys.argv:
    torch.set_num_threads(1)
  from spriteworld_b200 import environment, workloads
  wl = workloads.WORKLOADS['c2']()
  E = wl.n_envs
  if '--mallopt' in sys.argv:
    print('tune_host_allocator:', environment.tune_host_allocator(), flush=True)
  for procs, threads, K in ((0, 1, 32), (0, 1, 32)):
    env = environment.BatchedEnvironment(n_envs=E, device=0, rng=np.random.RandomState(1), pool_depth=K,
                                         refill_threads=threads, refill_procs=procs, **wl.plugin_config())
    acts = torch.from_numpy(wl.sample_actions(np.random.RandomState(5), 16, E)).to(env.engine.device)
    for i in range(60):
      env.step(acts[i % 16])
    torch.cuda.synchronize()
    n = 300
    c0 = cpu_stat()
    t0 = time.perf_counter()
    for i in range(n):
      env.step(acts[i % 16])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = env.refill_stats()
    c1 = cpu_stat()
    print('    cgroup cpu.stat delta:', {k: int(c1[k]) - int(c0[k]) for k in c1 if k in ('usage_usec', 'nr_periods', 'nr_throttled', 'throttled_usec')}, flush=True)
    import glob
    tt = []
    for t in glob.glob('/proc/self/task/*/stat'):
      try:
        f = open(t).read()
        comm = f[f.index('(') + 1:f.rindex(')')]
        rest = f[f.rindex(')') + 2:].split()
        tt.append((int(rest[11]) + int(rest[12]), comm))
      except Exception:
        pass
    tt.sort(reverse=True)
    print('    parent: %d threads, cpu ticks (10 ms) of the top ones: %s; sum %d' % (len(tt), tt[:10], sum(t[0] for t in tt)), flush=True)
    print('procs %d threads %d K %d: %.2f M env-steps/s, %.3f ms/step; refills %d scenes %d host %.3fs (%.2f us/scene) blocked %d %.3fs'
          % (procs, threads, K, E * n / dt / 1e6, 1e3 * dt / n, st['refills'], st['scenes'], st['host_seconds'],
             1e6 * st['host_seconds'] / max(st['scenes'], 1), st['blocked'], st['blocked_seconds']), flush=True)
    print('    plan %.3fs collect %.3fs upload %.3fs; inside the workers %.3fs; affinity %d cpus' % (
        st.get('plan_seconds', 0.0), st['sample_seconds'], st['upload_seconds'],
        env._pool.worker_seconds if env._pool is not None else 0.0, len(os.sched_getaffinity(0))), flush=True)
    env.close()


if __name__ == '__main__':
  main()
