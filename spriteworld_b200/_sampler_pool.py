"""Worker processes that sample and pack scenes for the batched environment's refill.

`init_sprites()` stays on the host (north_star), and in Python: a scene costs about 3 us to
draw and pack with NumPy, GIL-bound, so threads do not scale it.  At C2's reset rate (every env
every 20 steps) 4096 envs want ~1.5 M scenes/s; a few processes deliver that.  Each worker
holds the generator, the task filters and the colour map (sent once, cloudpickle: generators
are closures), receives (n, seed) and returns the packed scene arrays
(scene.arrays_from_layout) -- the same arrays the in-process path uploads.  Workers import
NumPy and this package's host modules only (no torch, no CUDA).
"""
import multiprocessing as mp
import threading

import numpy as np


def _worker_main(conn, payload):
  import cloudpickle
  from spriteworld_b200 import scene, sprite_generators
  init_sprites, n_slots, filters, color_to_rgb = cloudpickle.loads(payload)
  while True:
    try:
      msg = conn.recv()
    except EOFError:
      return
    if msg is None:
      return
    n, seed = msg
    try:
      np.random.seed(seed ^ 0x5BD1E995)    # generators whose callables draw from the global stream
      layout = sprite_generators.batch_of(init_sprites, n, np.random.RandomState(seed))
      conn.send(scene.arrays_from_layout(layout, n_slots, filters, color_to_rgb))
    except Exception as ex:   # reported to the caller, the worker lives on
      conn.send(ex)


class SamplerPool(object):

  def __init__(self, n_procs, init_sprites, n_slots, filters, color_to_rgb):
    import cloudpickle
    payload = cloudpickle.dumps((init_sprites, n_slots, list(filters), color_to_rgb))
    ctx = mp.get_context('spawn')
    self._conns, self._procs, self._locks = [], [], []
    for _ in range(int(n_procs)):
      parent, child = ctx.Pipe()
      p = ctx.Process(target=_worker_main, args=(child, payload), daemon=True)
      p.start()
      child.close()
      self._conns.append(parent)
      self._procs.append(p)
      self._locks.append(threading.Lock())

  def __len__(self):
    return len(self._procs)

  def sample(self, worker, n, seed):
    """Packed scene arrays of `n` scenes drawn with `seed`, from worker `worker` (blocking;
    call it from one thread per worker to keep them all busy)."""
    with self._locks[worker]:
      self._conns[worker].send((int(n), int(seed)))
      out = self._conns[worker].recv()
    if isinstance(out, Exception):
      raise out
    return out

  def close(self):
    for c in self._conns:
      try:
        c.send(None)
        c.close()
      except Exception:  # pragma: no cover
        pass
    for p in self._procs:
      p.join(timeout=2)
      if p.is_alive():  # pragma: no cover
        p.terminate()
    self._conns, self._procs = [], []
