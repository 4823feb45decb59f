"""Workload run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).

  compute-sanitizer --tool racecheck python tools/sanitize.py

Small batches of the BASELINE workloads (C2 5 sprites, C3 9 sprites, C5 128x128 with two
bands per frame and the Embodied action space) stepped past an auto-reset through the three
launch paths: swb_step_render (device buffers), swb_step_host (host buffers, chunked
renders) and swb_step_render_gather with two targets (render_kernel<true>, the variant that
also stores every frame into a second buffer: here a second buffer of the same device stands in
for a peer's).  The render kernel aliases its scratch area between phases and hands the staged
frame to the bulk-copy engine while the next frame's set-up runs: what racecheck is for.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  import torch
  from spriteworld_b200 import _native, workloads
  E = int(os.environ.get('SWB_SAN_ENVS', '24'))
  for key in ('c2', 'c3', 'c5'):
    wl = workloads.WORKLOADS[key]()
    T = 6
    eng, raster, _ = workloads.build_engine(wl, E, 4, device=0, seed=3, max_episode_length=3)
    acts = wl.sample_actions(np.random.RandomState(1), T, E)
    frames = raster.new_frames()
    for t in range(T):
      eng.step(torch.from_numpy(acts[t]).cuda(), raster, frames)
    for t in range(2):
      eng.step_host(acts[t], raster)
    # two targets on one device: the peer-store variant of the kernel
    H, W = wl.image_size[1], wl.image_size[0]
    bufs = [torch.zeros((2 * E, H, W, 3), dtype=torch.uint8, device='cuda') for _ in range(2)]
    ptrs = (ctypes.c_void_p * 2)(bufs[0].data_ptr(), bufs[1].data_ptr())
    for t in range(3):
      eng.step_gather(torch.from_numpy(acts[t]).cuda(), raster, (ptrs, 2, 0, bufs[0]))
    torch.cuda.synchronize()
    assert torch.equal(bufs[0][:E], bufs[1][:E]) and bufs[0][:E].any()
    print('%s: %d envs, %d kernel launches' % (wl.name, E, eng.launch_count()), flush=True)
    raster.close()
    eng.close()


if __name__ == '__main__':
  main()
