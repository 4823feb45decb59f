"""Benchmark of the hot path: env-steps/s (= rendered frames/s) of step+render.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5]
    python bench.py --impl reference ...      # the CPU path (oracle port) on the host cores

One "step" = one Environment.step (action -> pose/velocity update -> reward ->
termination/auto-reset -> PILRenderer frame) for every env of the batch.  Prints ONE JSON
line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'env_steps_per_sec'
UNIT = 'env-steps/s'


def _peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)['hbm_gbs']), 'measured'
  return 6650.0, 'fallback'


class ClockSampler(threading.Thread):
  """nvidia-smi SM clock / throttle-reason samples during the timed region."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.rows, self.proc = index, [], None

  def run(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
           '--format=csv,noheader,nounits', '-lms', '100'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      for line in self.proc.stdout:
        self.rows.append([c.strip() for c in line.split(',')])
    except Exception:
      pass

  def stop(self):
    if self.proc is not None:
      self.proc.terminate()
    self.join(timeout=2)
    sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
    reasons = set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows:
      for n, v in zip(names, r[3:7]):
        if v == 'Active':
          reasons.add(n)
    return dict(sm_mhz=float(np.median(sm)) if sm else None,
                sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                samples=len(sm))


def usable_cores():
  """Host CPUs this process may actually use: the affinity mask capped by the cgroup CPU
  quota (cpu.max).  On the GPU boxes 128 logical CPUs are visible but the container's quota
  is 16; more threads than that only add throttling."""
  try:
    n = len(os.sched_getaffinity(0))
  except AttributeError:
    n = os.cpu_count() or 1
  quota = None
  try:   # cgroup v2
    with open('/sys/fs/cgroup/cpu.max') as f:
      q, period = f.read().split()[:2]
    if q != 'max':
      quota = int(q) / float(period)
  except (OSError, ValueError):
    try:   # cgroup v1
      with open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us') as f:
        q = int(f.read())
      with open('/sys/fs/cgroup/cpu/cpu.cfs_period_us') as f:
        period = int(f.read())
      if q > 0 and period > 0:
        quota = q / float(period)
    except (OSError, ValueError):
      pass
  if quota:
    n = max(1, min(n, int(round(quota))))
  return n


def cpu_reference(wl, n_sample_envs, steps, warmup, seed=1000, cores=None):
  """The reference's algorithm on the host cores: oracle port (C restatement of the
  reference path incl. Pillow's polygon fill and LANCZOS), one thread per core, each
  stepping its own slice of a bounded env sample.  Returns (env-steps/s, cores, seconds)."""
  from concurrent.futures import ThreadPoolExecutor
  from oracle import oracle
  from spriteworld_b200 import constants
  from tests import fixtures
  cores = cores or usable_cores()
  n = max(cores, (n_sample_envs // cores) * cores)
  K = max(2, (steps + warmup) // wl.max_episode_length + 2)
  rng = np.random.RandomState(seed)
  scenes = wl.sample_scenes(rng, n * K)
  rec = np.zeros((n * K, wl.n_slots), oracle.SPRITE_DTYPE)
  for f in ('x', 'y', 'm00', 'm01', 'm10', 'm11', 'vx', 'vy', 'member', 'shape', 'pos_f32', 'rgb'):
    rec[f] = scenes[f]
  pool = rec.reshape(n, K, wl.n_slots)
  cfg = fixtures.env_cfg_from_meta(dict(
      action=wl.action, keep_in_frame=True, max_episode_length=wl.max_episode_length,
      nodes=[dict(n, goal=list(n.get('goal', (0, 0))), weights=list(n.get('weights', (1, 1))))
             if n['kind'] == 'find_goal' else n for n in wl.nodes]))
  tab = oracle.shape_table(constants.SHAPES)
  rc = oracle.raster_cfg(wl.image_size[0], wl.image_size[1], wl.anti_aliasing)
  bo = oracle.BatchOracle(cfg, tab, rc, pool)
  actions = wl.sample_actions(np.random.RandomState(7), steps + warmup, n)
  chunk = n // cores
  with ThreadPoolExecutor(cores) as ex:
    def one_step(t):
      list(ex.map(lambda c: bo.step(actions[t], c * chunk, (c + 1) * chunk), range(cores)))
    for t in range(warmup):
      one_step(t)
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
      one_step(t)
    dt = time.perf_counter() - t0
  return n * steps / dt, cores, dt, n


def run_reference_arm(args, wl):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  # each step = a bounded sample of the workload sized to finish within minutes
  n_sample = 128 * usable_cores()
  value, cores, dt, n = cpu_reference(wl, n_sample, args.steps, args.warmup)
  sample = ('%d envs x %d steps of %s on %d host threads = usable cores (cgroup quota; %d logical '
            'CPUs visible), oracle C port of the reference path' % (
                n, args.steps, wl.name, cores, os.cpu_count() or 1))
  line = dict(
      impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus,
      steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps,
      higher_is_better=True, scaling='weak', vs_baseline=None, dtype='u8', data='synthetic',
      config=dict(workload=wl.name, sample_envs=n),
      cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind='port', sample=sample),
      e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=20)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--workload', default='c2', choices=['c2', 'c3', 'c4', 'c5'])
  ap.add_argument('--envs', type=int, default=0, help='override envs per GPU')
  ap.add_argument('--gather', default='peer', choices=['peer', 'ce', 'nccl'],
                  help='N>1 frame gather: stores into peer memory from the render kernel, or a '
                       'separate NCCL all-gather')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  args = ap.parse_args()
  if args.warmup < 3:
    args.warmup = 3
  from spriteworld_b200 import workloads
  wl = workloads.WORKLOADS[args.workload]()
  if args.impl == 'reference':
    return run_reference_arm(args, wl)

  import torch
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)

  E = args.envs or wl.n_envs           # per GPU: weak scaling by env index
  T = args.warmup + args.steps
  K = T // wl.max_episode_length + 3   # pooled scenes per env cover every auto-reset
  eng, raster, _ = workloads.build_engine(wl, E, K, device=local_rank, seed=1000 + rank)
  actions = torch.from_numpy(wl.sample_actions(np.random.RandomState(7 + rank), T, E)).to(dev)
  H, W = wl.image_size[1], wl.image_size[0]
  frame_bytes = E * H * W * 3
  # frame ring larger than L2 (126 MB) so that every step's frame writes reach HBM
  n_ring = max(2, int(np.ceil(160e6 / frame_bytes)) + 1)
  ring = [raster.new_frames() for _ in range(n_ring)]
  gathered, peer, inflight, n_gslots = None, None, [], 2   # inflight: (work handle, ring slot read, gather slot)
  if world > 1 and args.gather == 'nccl':   # double-buffered destination of the per-step all-gather
    gathered = [torch.empty((world * E, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
  elif world > 1:
    from spriteworld_b200 import _native, distributed
    # as many gathered buffers as make one pass over them larger than L2
    n_gslots = max(2, int(np.ceil(160e6 / (world * frame_bytes))) + 1)
    try:
      peer = distributed.PeerFrames(E, (H, W, 3), dev, n_slots=n_gslots)
    except _native.NativeError as ex:   # raised on every rank or on none
      if rank == 0:
        sys.stderr.write('peer-memory gather unavailable (%s); using the NCCL all-gather\n' % ex)
      args.gather = 'nccl'
      gathered = [torch.empty((world * E, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]

  def wait_for(pred):
    for item in list(inflight):
      if pred(item):
        item[0].wait()
        inflight.remove(item)

  def one_step(t, gather=True):
    slot, dst = t % n_ring, t % (n_gslots if peer is not None else 2)
    if world > 1 and gather and peer is not None:
      # the single collective of the path, fused: the render kernel stores each frame into
      # every rank's gathered buffer over NVLink; a one-element all-reduce on NCCL's stream
      # is the completion barrier and overlaps the next step
      wait_for(lambda it: it[2] == dst)      # everyone is done with the step that last used dst
      if args.gather == 'ce':
        # variant: render into this rank's block, then copy-engine pushes to the peers
        eng.step(actions[t % T], raster, peer.own_slab(dst))
        inflight.append((peer.push(dst), -1, dst))
      else:
        eng.step_gather(actions[t % T], raster, peer.slot(dst))
        inflight.append((peer.barrier(async_op=True), -1, dst))
      return
    wait_for(lambda it: it[1] == slot)     # the gather that last read this ring buffer
    fr = ring[slot]
    eng.step(actions[t % T], raster, fr)
    if world > 1 and gather:
      # the single collective of the path as a separate NCCL all-gather.  It runs on NCCL's
      # stream and overlaps the next step's compute.
      wait_for(lambda it: it[2] == dst)
      inflight.append((dist.all_gather_into_tensor(gathered[dst], fr, async_op=True), slot, dst))

  def drain():
    wait_for(lambda it: True)

  def barrier():
    drain()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for t in range(args.warmup):
    one_step(t)
  barrier()
  sampler = ClockSampler(local_rank) if rank == 0 else None
  if sampler:
    sampler.start()
    time.sleep(0.3)
  launches0 = eng.launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  ev0.record()
  for t in range(args.warmup, T):
    one_step(t)
  drain()
  ev1.record()
  barrier()
  ms = ev0.elapsed_time(ev1)
  launches = eng.launch_count() - launches0
  if world > 1:
    tms = torch.tensor([ms], device=dev)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())

  # SURVEY 8(e) asks for both numbers: the same steps with the frames left sharded
  sharded = None
  if world > 1:
    evs0, evs1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    evs0.record()
    for t in range(T, T + args.steps):
      one_step(t, gather=False)
    evs1.record()
    barrier()
    tms = torch.tensor([evs0.elapsed_time(evs1)], device=dev)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    sharded = dict(value=world * E * args.steps / (float(tms.item()) * 1e-3), unit=UNIT,
                   ms_per_step=float(tms.item()) / args.steps)

  # dominant kernel alone: K launches of the render kernel, CUDA events on its stream
  evr0, evr1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n_r = max(20, min(args.steps, 100))
  for i in range(3):
    eng.render(raster, ring[i % n_ring])
  torch.cuda.synchronize()
  evr0.record()
  for i in range(n_r):
    eng.render(raster, ring[i % n_ring])
  evr1.record()
  torch.cuda.synchronize()
  render_ms = evr0.elapsed_time(evr1) / n_r
  clocks = sampler.stop() if sampler else None

  # end to end through the C-ABI with HOST buffers (pinned): H2D actions, D2H frames+outputs
  e2e = None
  if not args.no_e2e:
    a_host = torch.from_numpy(wl.sample_actions(np.random.RandomState(99 + rank), 8, E)).pin_memory()
    out = dict(
        frames=torch.empty((E, H, W, 3), dtype=torch.uint8).pin_memory().numpy(),
        reward=torch.empty(E, dtype=torch.float64).pin_memory().numpy(),
        step_type=torch.empty(E, dtype=torch.int8).pin_memory().numpy(),
        success=torch.empty(E, dtype=torch.uint8).pin_memory().numpy(),
        status=torch.empty(E, dtype=torch.uint8).pin_memory().numpy())
    n_e2e = max(10, min(args.steps, 50))
    a_np = a_host.numpy()
    for i in range(3):
      eng.step_host(a_np[i % 8], raster, out=out)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
      eng.step_host(a_np[i % 8], raster, out=out)
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
      tdt = torch.tensor([dt], device=dev)
      dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
      dt = float(tdt.item())
    e2e = dict(value=world * E * n_e2e / dt, unit=UNIT,
               h2d_bytes_per_step=int(a_np[0].nbytes),
               d2h_bytes_per_step=int(frame_bytes + E * (8 + 1 + 1 + 1)),
               steps=n_e2e, ms_per_step=1e3 * dt / n_e2e,
               path='swb_step_host: pinned host actions -> H2D -> step+render -> D2H frames, '
                    'reward, step_type, success, status -> stream sync')

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  peak, peak_kind = _peaks()
  alg_bytes = wl.algorithmic_bytes() * E
  achieved = alg_bytes / (render_ms * 1e-3) / 1e9
  traffic = None
  tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
  if os.path.exists(tpath):
    with open(tpath) as f:
      traffic = json.load(f).get(args.workload)
  cpu = None
  if not args.no_cpu_baseline:
    n_cpu_steps = 25
    v, cores, dt, n = cpu_reference(wl, 512 * usable_cores(), n_cpu_steps, 3)
    v1 = cpu_reference(wl, 32, 10, 2, cores=1)[0]
    cpu = dict(value=v, unit=UNIT, cores=cores, kind='port', single_core=v1,
               sample='%d envs x %d steps of %s, %.1f s wall, oracle C port of the reference '
                      'path (Pillow polygon fill + LANCZOS restated), one thread per usable core '
                      '(affinity mask capped by the cgroup CPU quota; %d logical CPUs visible)'
                      % (n, n_cpu_steps, wl.name, dt, os.cpu_count() or 1))
  value = world * E * args.steps / (ms * 1e-3)
  line = dict(
      metric=METRIC, value=value, unit=UNIT, frames_per_sec=value, n_gpus=world,
      steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
      scaling='weak', vs_baseline=None, dtype='u8', data='synthetic',
      config=dict(workload=wl.name, envs_per_gpu=E, n_sprites=wl.n_slots,
                  image=[H, W, 3], anti_aliasing=wl.anti_aliasing,
                  max_episode_length=wl.max_episode_length, auto_reset='pooled scenes',
                  pool_depth=K, l2='frame ring of %d buffers (%.0f MB) > L2, no flush'
                  % (n_ring, n_ring * frame_bytes / 1e6),
                  collective=('none' if world == 1 else
                              'frames stored into every rank\'s gathered buffer by the render kernel '
                              '(NVLink peer memory) + one-element NCCL all-reduce as completion barrier'
                              if peer is not None and args.gather == 'peer' else
                              'render into the rank\'s block of the gathered buffer, copy-engine pushes to '
                              'the peers over NVLink + one-element NCCL all-reduce as completion barrier'
                              if peer is not None else
                              'NCCL all_gather of frames per step (async, overlaps the next step)')),
      roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s',
                    frac=achieved / peak, traffic=traffic, peak_kind=peak_kind,
                    kernel='render_kernel', kernel_ms=render_ms,
                    algorithmic_bytes_per_launch=alg_bytes),
      cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks)
  if sharded:
    line['frames_sharded'] = sharded   # same steps without the frame gather
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
