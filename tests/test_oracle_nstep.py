"""Pins oracle.nstep against the reference's 7 golden adder cases
(acme/adders/reverb/transition_test.py:29-170; harness semantics test_utils.py:133-224)."""
import json
import os

import numpy as np
import pytest

from acme_b200 import dm_env
from oracle import nstep

CASES = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'nstep_cases.json')))


def make_timestep(step):
  if step['kind'] == 'mid':
    return dm_env.transition(reward=step['reward'], observation=step['observation'], discount=step['discount'])
  return dm_env.termination(reward=step['reward'], observation=step['observation'])


def assert_item_close(expected, observed):
  assert len(expected) == len(observed)
  for e, o in zip(expected, observed):
    if isinstance(e, dict):
      assert set(e) == set(o)
      for k in e:
        np.testing.assert_array_almost_equal(e[k], o[k])
    else:
      np.testing.assert_array_almost_equal(e, o)


@pytest.mark.parametrize('case', CASES, ids=[c['name'] for c in CASES])
def test_reference_golden_cases(case):
  client = nstep.RecordingClient()
  adder = nstep.ReferenceAdder(client, case['n_step'], case['additional_discount'])
  steps = case['steps']
  adder.add_first(dm_env.restart(case['first']))
  for s in steps[:-1]:
    adder.add(0, make_timestep(s), extras=s['extras'] or ())
  if len(steps) == 1:
    assert not client.writers
  else:
    assert len(client.writers) == 1 and not client.writers[0].closed
  adder.add(0, make_timestep(steps[-1]), extras=steps[-1]['extras'] or ())
  # episode end closes the (lazily created) writer; no new one yet
  assert len(client.writers) == 1 and client.writers[0].closed
  observed = [p[1] for p in client.writers[0].priorities]
  assert len(observed) == len(case['expected'])
  for e, o in zip(case['expected'], observed):
    assert_item_close(e, o)
  assert all(p[0] == nstep.DEFAULT_TABLE and p[2] == 1.0 for p in client.writers[0].priorities)
  # a second trajectory gets a fresh writer
  adder.add_first(dm_env.restart(case['first']))
  adder.add(0, make_timestep(steps[0]), extras=steps[0]['extras'] or ())
  assert len(client.writers) == 2
  assert client.writers[1].closed == (steps[0]['kind'] == 'term')


def test_protocol_errors():
  adder = nstep.ReferenceAdder(nstep.RecordingClient(), 2, 0.9)
  with pytest.raises(ValueError):
    adder.add(0, dm_env.transition(0., 1))
  with pytest.raises(ValueError):
    adder.add_first(dm_env.transition(0., 1))
  adder.add_first(dm_env.restart(0))
  with pytest.raises(ValueError):
    adder.add_first(dm_env.restart(0))


@pytest.mark.parametrize('case', CASES, ids=[c['name'] for c in CASES])
def test_enumeration_matches_object_form(case):
  """App. A.1 closed form == what the state machine emits (start obs, arrival obs, window length)."""
  if isinstance(case['first'], dict):
    pytest.skip('dict observations')
  T, n = len(case['steps']), case['n_step']
  obs = [case['first']] + [s['observation'] for s in case['steps']]
  rew = [np.float32(s['reward']) for s in case['steps']]
  dis = [np.float32(s['discount']) for s in case['steps']]
  items = nstep.enumerate_items(T, n)
  assert len(items) == T + min(n, T) - 1 == len(case['expected'])
  for (start, length), exp in zip(items, case['expected']):
    R, D = nstep.nstep_return(rew[start:start + length], dis[start:start + length], case['additional_discount'])
    assert R.dtype == np.float32 and D.dtype == np.float32
    np.testing.assert_array_almost_equal([obs[start], R, D, obs[start + length]], [exp[0], exp[2], exp[3], exp[4]])


def test_fp32_arithmetic_is_unfused():
  rng = np.random.default_rng(0)
  r = rng.standard_normal(5).astype(np.float32)
  d = rng.uniform(0.5, 1, 5).astype(np.float32)
  g = np.float32(0.99)
  R, D = nstep.nstep_return(r, d, g)
  Rm, Dm = r[0], d[0]
  for j in range(1, 5):
    Dm = np.float32(Dm * g)
    Rm = np.float32(Rm + np.float32(r[j] * Dm))
    Dm = np.float32(Dm * d[j])
  assert R == Rm and D == Dm
