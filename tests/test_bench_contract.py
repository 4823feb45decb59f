"""bench.py's reference arm runs on the host: check the JSON line it prints (CPU tier)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--steps', '1', '--warmup', '3'], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d['impl'] == 'reference' and d['metric'] == 'env_steps_per_sec'
  assert d['unit'] == 'env-steps/s' and d['higher_is_better'] is True and d['value'] > 0
  assert d['config']['workload'] == 'goal_finding_select_move_4096x5_64x64'
  assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
  assert d['cpu_baseline']['value'] == d['value']
  assert d['e2e'] == dict(value=d['value'], unit=d['unit'], h2d_bytes_per_step=0,
                          d2h_bytes_per_step=0)


def test_other_ranks_of_the_reference_arm_exit_quietly():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--gpus', '2', '--steps', '1', '--warmup', '3'], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=120)
  assert out.returncode == 0 and out.stdout.strip() == ''


def test_usable_cores_respects_affinity():
  sys.path.insert(0, ROOT)
  import bench
  n = bench.usable_cores()
  assert 1 <= n <= len(os.sched_getaffinity(0))
