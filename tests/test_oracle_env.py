"""Pins the oracle's Environment.step restatement (oracle/sw_env_oracle.c) against
episodes stepped by the REAL reference (tests/golden/episodes_*.npz).

Contract: positions, step_type, success bit-exact; frames bit-exact;
FindGoalPosition / MetaAggregated rewards bit-exact (the oracle calls the same libm pow);
Clustering rewards to 1e-6 relative (scikit-learn's float32 distance blocks).
"""
import numpy as np
import pytest

from oracle import oracle
from tests import fixtures

CASES = ['goal_finding', 'more_targets_f64', 'clustering', 'sorting', 'embodied', 'moving']


def _run(name):
  ep = fixtures.Episodes(name)
  cfg, tab, rc, pool = ep.oracle_parts()
  bo = oracle.BatchOracle(cfg, tab, rc, pool)
  out = dict(pos=[], reward=[], step_type=[], success=[], frames=[], cursor=[])
  for t in range(ep.T):
    bo.step(ep.actions[t])
    out['pos'].append(np.stack([bo.cur['x'], bo.cur['y']], -1))
    out['reward'].append(bo.reward.copy())
    out['step_type'].append(bo.step_type.copy())
    out['success'].append(bo.success.copy())
    out['cursor'].append(bo.cursor.copy())
    out['frames'].append(bo.frames[ep.meta['frame_envs']].copy())
  return ep, {k: np.stack(v) for k, v in out.items()}


@pytest.mark.parametrize('name', CASES)
def test_episode_matches_reference(name):
  ep, got = _run(name)
  assert np.array_equal(got['step_type'], ep.step_type)
  assert np.array_equal(got['cursor'], ep.scene_idx)
  assert np.array_equal(got['success'], ep.success)
  occupied = ep.scenes['shape'][np.arange(ep.E)[None, :], ep.scene_idx] > 0   # (T, E, S)
  assert np.array_equal(got['pos'][occupied], ep.pos[occupied])               # bit-exact
  mid = ep.step_type != 0
  if name == 'clustering':
    np.testing.assert_allclose(got['reward'][mid], ep.reward[mid], rtol=1e-6, atol=1e-9)
  else:
    assert np.array_equal(got['reward'][mid], ep.reward[mid], equal_nan=True)
  assert np.array_equal(got['frames'], ep.frames)
  assert (ep.step_type == 2).sum() > 0 and (ep.step_type == 0).sum() > ep.E


def test_fixtures_exercise_the_interesting_paths():
  ep = fixtures.Episodes('goal_finding')
  moved = np.abs(np.diff(ep.pos, axis=0)).sum(axis=(2, 3)) > 0
  assert moved.sum() > 100                       # clicks do hit sprites
  ep = fixtures.Episodes('embodied')
  d = np.abs(np.diff(ep.pos, axis=0)) > 0
  assert (d[:, :, :-1].any(axis=(2, 3))).sum() > 5   # something was carried
  ep = fixtures.Episodes('moving')
  assert np.isnan(ep.reward).sum() == 0          # nanmean skips the empty-filter subtask
  assert (ep.pos < 0).any() or (ep.pos > 1).any()    # sprites leave the frame
