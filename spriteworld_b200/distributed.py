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
