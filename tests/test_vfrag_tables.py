"""The vertical pass's tensor-core tables against the direct LANCZOS convolution (CPU tier).

swb_tables.h cuts Pillow's 22-bit vertical taps into three int8 limbs and lays them out as the
B fragments of mma.sync.m16n8k32 per block of eight output rows.  tests/native/check_vfrag.cpp
expands the fragments back through the instruction's register layout on the host and compares
d0 + 2^8 d1 + 2^16 d2 with the plain convolution for several (canvas, image) sizes, including
non-multiples of eight and anti_aliasing 1..5; the GPU tier then checks whole frames."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_include():
  for d in (os.environ.get('CUDA_HOME'), '/usr/local/cuda'):
    if d and os.path.exists(os.path.join(d, 'include', 'cuda_runtime.h')):
      return os.path.join(d, 'include')
  return None


def test_vertical_fragments_reproduce_the_taps(tmp_path):
  cxx, inc = shutil.which('g++'), _cuda_include()
  if not cxx or not inc:
    pytest.skip('needs g++ and the CUDA headers')
  exe = str(tmp_path / 'check_vfrag')
  src = os.path.join(ROOT, 'tests', 'native', 'check_vfrag.cpp')
  build = subprocess.run([cxx, '-std=c++17', '-O1', '-I', inc, '-o', exe, src], capture_output=True, text=True)
  assert build.returncode == 0, build.stderr[-2000:]
  run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
  assert run.returncode == 0, run.stdout + run.stderr
  lines = [l for l in run.stdout.splitlines() if '->' in l]
  assert len(lines) >= 10 and all(l.endswith('mismatches 0') for l in lines), run.stdout
  # 5x reduction: tap 29 of the interior vector is zero, so a block needs two k-steps only
  assert any(l.startswith('320 -> 64: k-steps 2, classes 3') for l in lines), run.stdout
