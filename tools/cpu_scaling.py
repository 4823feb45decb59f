"""How the CPU reference arm (oracle port) scales with host threads on this box.

  python tools/cpu_scaling.py [--threads 1,8,32,64,128]

bench.py's cpu_baseline / --impl reference use one thread per core; this prints env-steps/s
per thread count so that the choice can be checked against the box (SMT siblings, memory
bandwidth and the allocator all bend the curve).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--threads', default='1,8,32,64,128')
  ap.add_argument('--workload', default='c2')
  ap.add_argument('--steps', type=int, default=10)
  args = ap.parse_args()
  import bench
  from spriteworld_b200 import workloads
  wl = workloads.WORKLOADS[args.workload]()
  print('host cpus: %s, usable (affinity and cgroup quota): %d' % (os.cpu_count(), bench.usable_cores()))
  for c in [int(x) for x in args.threads.split(',')]:
    v, cores, dt, n = bench.cpu_reference(wl, 16 * c, args.steps, 3, cores=c)
    print('threads %4d: %9.0f env-steps/s  (%.2f ms per env-step per thread, %d envs, %.1f s)' % (
        c, v, 1e3 * c / v, n, dt))


if __name__ == '__main__':
  main()
