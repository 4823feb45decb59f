"""Multi-GPU plumbing: shard envs by index, one gather of the rendered frames per step.

Envs are independent (no term of the step path couples two envs), so the batch is cut into
contiguous blocks [start, start+count) per rank and each rank owns one engine.  The only
collective of the path is the gather of the step's outputs (frames dominate: H*W*3 bytes
per env) over NCCL/NVLink; rewards and flags ride along.  `torch.distributed` is used for
the plumbing only; the same code runs on gloo/CPU tensors for the tests.
"""
import collections

import torch
import torch.distributed as dist


def env_shard(n_envs_total, rank, world_size):
  """Contiguous block of env indices owned by `rank`: (start, count)."""
  base, extra = divmod(int(n_envs_total), int(world_size))
  count = base + (1 if rank < extra else 0)
  start = rank * base + min(rank, extra)
  return start, count


def shard_sizes(n_envs_total, world_size):
  return [env_shard(n_envs_total, r, world_size)[1] for r in range(world_size)]


GatheredStep = collections.namedtuple('GatheredStep', ['frames', 'reward', 'step_type', 'success'])


class StepGatherer(object):
  """Preallocated all-gather of per-rank step outputs into global (env-ordered) tensors."""

  def __init__(self, n_envs_total, frame_shape, device, group=None):
    self.group = group
    self.world = dist.get_world_size(group)
    self.rank = dist.get_rank(group)
    self.sizes = shard_sizes(n_envs_total, self.world)
    self.equal = len(set(self.sizes)) == 1
    E = int(n_envs_total)
    self.frames = torch.empty((E,) + tuple(frame_shape), dtype=torch.uint8, device=device)
    self.reward = torch.empty(E, dtype=torch.float64, device=device)
    self.step_type = torch.empty(E, dtype=torch.int8, device=device)
    self.success = torch.empty(E, dtype=torch.uint8, device=device)

  def _gather(self, out, local):
    local = local.contiguous()
    if self.equal:
      dist.all_gather_into_tensor(out, local, group=self.group)
      return
    # uneven shards: collectives want equal sizes, so pad every rank's block to the largest
    big = max(self.sizes)
    padded = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((self.world * big,) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=self.group)
    start = 0
    for r, n in enumerate(self.sizes):
      out[start:start + n] = buf[r * big:r * big + n]
      start += n

  def gather(self, frames, reward, step_type, success):
    """Every rank receives every env's outputs, ordered by global env index."""
    self._gather(self.frames, frames)
    self._gather(self.reward, reward)
    self._gather(self.step_type, step_type)
    self._gather(self.success, success)
    return GatheredStep(self.frames, self.reward, self.step_type, self.success)


def _native_error(msg):
  from spriteworld_b200 import _native
  return _native.NativeError('PeerFrames: ' + msg)


class PeerFrames(object):
  """Gathered frame buffers every rank's render kernel stores into over NVLink peer memory.

  The frame gather is the path's one collective.  Instead of running it as a
  separate NCCL all-gather after the render kernel -- whose copy CTAs then compete with the
  next step's render for SMs -- each rank maps the other ranks' gathered buffers (CUDA IPC)
  and the render kernel's write-out phase stores every finished frame into all of them
  (`swb_step_render_gather`).  What remains of the collective is a completion barrier: a
  one-element all-reduce enqueued after the kernel, on NCCL; with `host_barrier=True`
  (tests with several ranks on one device, where NCCL refuses to run) a device
  synchronise plus a host barrier on the default group.

  `n_slots` buffers are kept so that a consumer can still read step t while step t+1 is
  being written.
  """

  def __init__(self, n_envs_per_rank, frame_shape, device, n_slots=2, group=None,
               host_barrier=False):
    import ctypes
    import numpy as np
    from spriteworld_b200 import _native
    from spriteworld_b200 import engine as engine_lib
    self._lib = _native.load()
    self.group = group
    self.world = dist.get_world_size(group)
    self.rank = dist.get_rank(group)
    if self.world > _native.MAX_PEERS:
      raise ValueError('PeerFrames supports up to %d ranks' % _native.MAX_PEERS)
    self.device = torch.device(device)
    self.host_barrier = host_barrier
    self.E = int(n_envs_per_rank)
    self.n_slots = int(n_slots)
    dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
    shape = (self.world * self.E,) + tuple(frame_shape)
    nbytes = int(np.prod(shape))
    # my own buffers, exported.  Failures are agreed on by all ranks (a rank that raised
    # alone would leave the others waiting in the next collective).
    self._own, self._opened, self._ptrs, handles, error = [], [], [], [], None
    try:
      for _ in range(self.n_slots):
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * _native.IPC_HANDLE_BYTES)()
        _native.check(self._lib.swb_ipc_alloc(dev_index, nbytes, ctypes.byref(ptr), handle))
        self._own.append(ptr)
        handles.append(bytes(handle))
    except _native.NativeError as ex:
      error = 'rank %d: %s' % (self.rank, ex)
    everyone = [None] * self.world
    dist.all_gather_object(everyone, (error, handles), group=group)
    self._agree([e for e, _ in everyone])
    # the peers' buffers, mapped (per slot: ctypes array of world pointers, rank-ordered)
    try:
      for slot in range(self.n_slots):
        arr = (ctypes.c_void_p * self.world)()
        for r, (_, hs) in enumerate(everyone):
          if r == self.rank:
            arr[r] = self._own[slot].value
          else:
            p = ctypes.c_void_p()
            h = (ctypes.c_uint8 * _native.IPC_HANDLE_BYTES).from_buffer_copy(hs[slot])
            _native.check(self._lib.swb_ipc_open(dev_index, h, ctypes.byref(p)))
            self._opened.append(p)
            arr[r] = p.value
        self._ptrs.append(arr)
    except _native.NativeError as ex:
      error = 'rank %d: %s' % (self.rank, ex)
    errors = [None] * self.world
    dist.all_gather_object(errors, error, group=group)
    self._agree(errors)
    self.frames = [torch.as_tensor(engine_lib._DevicePointer(self._own[s].value, shape, '|u1'),
                                   device=self.device) for s in range(self.n_slots)]
    self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
    # copy-engine variant: one side stream per peer
    self._slab_bytes = nbytes // self.world
    self._copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.world)]
    self._sync_stream = torch.cuda.Stream(device=self.device)
    self._slab_read = [None] * self.n_slots   # event: the pushes that read my slab of slot s

  def _agree(self, errors):
    errors = [e for e in errors if e]
    if errors:
      self.close()
      raise _native_error('; '.join(errors))

  def slot(self, i):
    """Targets of step i for Engine.step_gather: (pointer list, n, env offset, local view)."""
    s = i % self.n_slots
    return self._ptrs[s], self.world, self.rank * self.E, self.frames[s]

  def own_slab(self, i):
    """This rank's block of gathered buffer i % n_slots: render straight into it, then
    push().  Waits (on the current stream) for the pushes that last read it."""
    s = i % self.n_slots
    if self._slab_read[s] is not None:
      torch.cuda.current_stream(self.device).wait_event(self._slab_read[s])
    return self.frames[s][self.rank * self.E:(self.rank + 1) * self.E]

  def push(self, i, async_op=True):
    """Copy-engine variant of the gather: after the render that filled own_slab(i), one
    peer-to-peer copy per rank on its own stream (DMA over NVLink, no SM involved), then
    the completion barrier on a side stream.  The caller's stream is not held up."""
    import ctypes
    from spriteworld_b200 import _native
    s = i % self.n_slots
    lo, hi = self.rank * self.E, (self.rank + 1) * self.E
    src = self.frames[s][lo:hi]
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(self.device))
    done = []
    for r in range(self.world):
      if r == self.rank:
        continue
      st = self._copy_streams[r]
      st.wait_event(ready)
      _native.check(self._lib.swb_peer_copy(
          ctypes.c_void_p(self._ptrs[s][r] + self.rank * self._slab_bytes),
          ctypes.c_void_p(src.data_ptr()), self._slab_bytes, ctypes.c_void_p(st.cuda_stream)))
      ev = torch.cuda.Event()
      ev.record(st)
      done.append(ev)
    for ev in done:
      self._sync_stream.wait_event(ev)
    with torch.cuda.stream(self._sync_stream):
      read = torch.cuda.Event()
      read.record(self._sync_stream)
      self._slab_read[s] = read
      if self.host_barrier:
        self._sync_stream.synchronize()
        dist.barrier(group=self.group)
        return None
      return dist.all_reduce(self._token, group=self.group, async_op=async_op)

  def barrier(self, async_op=False):
    """After it, the frames every rank stored for the steps enqueued so far are in place."""
    if self.host_barrier:
      torch.cuda.synchronize(self.device)
      dist.barrier(group=self.group)
      return None
    return dist.all_reduce(self._token, group=self.group, async_op=async_op)

  def close(self):
    for p in getattr(self, '_opened', []):
      self._lib.swb_ipc_close(p)
    self._opened = []
    self.frames = []
    for p in getattr(self, '_own', []):
      self._lib.swb_ipc_free(p)
    self._own = []
