"""Per-phase cycle breakdown of the render kernel (debug build, not the shipped library).

  python tools/phase_clocks.py [--workload c2]

Builds csrc/libspriteworld_b200_dbg.so with -DSWB_PHASE_CLOCKS (thread 0 of every CTA sums
clock64() deltas between the kernel's barriers into a device array), runs a few steps of the
workload and prints each phase's share of the CTA lifetime.  The shares are residency-weighted
wall time of a CTA under contention from its neighbours on the SM -- what ncu's instruction
counts cannot show (phases that keep one warp busy while seven wait).
"""
import argparse
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# index = id of the SWB_MARK that ends the phase (ids 0-2 are unused since the set-up phases
# were fused into one)
NAMES = ['-', '-', '-', 'A set-up (+tables)', 'B partition', 'B1 scan', 'B2 fold', 'background',
         'C H pass', 'C V pass', 'D write-out']


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--workload', default='c2')
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--no-build', action='store_true')
  args = ap.parse_args()
  from spriteworld_b200 import build as b
  dbg = os.path.join(b.CSRC, 'libspriteworld_b200_dbg.so')
  if not args.no_build:
    cmd = [b._nvcc()] + b.NVCC_FLAGS + ['-DSWB_PHASE_CLOCKS', '-o', dbg] + b.SOURCES
    subprocess.run(cmd, cwd=b.CSRC, check=True)
  if args.steps == 0:
    return
  from spriteworld_b200 import _native
  _native._LIB_PATH = dbg
  import torch
  from spriteworld_b200 import workloads
  import numpy as np
  wl = workloads.WORKLOADS[args.workload]()
  eng, raster, _ = workloads.build_engine(wl, wl.n_envs, 8, device=0)
  L = _native.load()
  L.swb_debug_phase_clocks.argtypes = [ctypes.c_void_p]
  acts = torch.from_numpy(wl.sample_actions(np.random.RandomState(7), args.steps + 3, wl.n_envs)).cuda()
  frames = raster.new_frames()
  out = (ctypes.c_ulonglong * 16)()
  for i in range(3):
    eng.step(acts[i], raster=raster, frames=frames)
  torch.cuda.synchronize()
  L.swb_debug_phase_clocks(out)
  for i in range(args.steps):
    eng.step(acts[3 + i], raster=raster, frames=frames)
  torch.cuda.synchronize()
  L.swb_debug_phase_clocks(out)
  tot = float(sum(out[:11]))
  for i, n in enumerate(NAMES):
    if n == '-':
      continue
    print('%-16s %6.2f%%  %8.0f cycles/CTA' % (n, 100.0 * out[i] / tot, out[i] / (args.steps * wl.n_envs)))
  print('total %.0f cycles/CTA' % (tot / (args.steps * wl.n_envs)))


if __name__ == '__main__':
  main()
