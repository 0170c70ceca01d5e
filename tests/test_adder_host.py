"""Host logic of the product adder without a GPU: protocol errors, writer life-cycle and the item
windows it asks the replay for (reference harness test_utils.py:179-224)."""
import json
import os

import numpy as np
import pytest

from acme_b200 import adders, dm_env
from oracle import nstep as onstep

CASES = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'nstep_cases.json')))


class FakeWriter:

  def __init__(self, *a, **k):
    self.steps, self.items, self.closed = [], [], False

  def append_step(self, observation, action, reward, discount, next_observation, extras=(), tables=None):
    assert not self.closed
    self.steps.append((observation, action, reward, discount, next_observation, extras))

  def create_item(self, table, num_timesteps, priority):
    assert not self.closed
    assert 1 <= num_timesteps <= len(self.steps)
    self.items.append((table, len(self.steps) - num_timesteps, num_timesteps, priority))

  def close(self):
    assert not self.closed
    self.closed = True


class FakeClient:

  def __init__(self):
    self.writers = []

  def writer(self, max_sequence_length, delta_encoded=False, chunk_length=None):
    self.writers.append(FakeWriter())
    return self.writers[-1]


def ts(step):
  if step['kind'] == 'mid':
    return dm_env.transition(reward=step['reward'], observation=step['observation'], discount=step['discount'])
  return dm_env.termination(reward=step['reward'], observation=step['observation'])


@pytest.mark.parametrize('case', CASES, ids=[c['name'] for c in CASES])
def test_windows_and_writer_lifecycle(case):
  client = FakeClient()
  adder = adders.NStepTransitionAdder(client, case['n_step'], case['additional_discount'])
  steps = case['steps']
  adder.add_first(dm_env.restart(case['first']))
  for s in steps[:-1]:
    adder.add(0, ts(s), extras=s['extras'] or ())
  if len(steps) == 1:
    assert not client.writers
  else:
    assert len(client.writers) == 1 and not client.writers[0].closed
  adder.add(0, ts(steps[-1]), extras=steps[-1]['extras'] or ())
  assert len(client.writers) == 1 and client.writers[0].closed
  w = client.writers[0]
  assert [(s, l) for _, s, l, _ in w.items] == onstep.enumerate_items(len(steps), case['n_step'])
  assert all(t == adders.DEFAULT_PRIORITY_TABLE and p == 1. for t, _, _, p in w.items)
  # materialising the windows on the host reproduces the reference's golden transitions
  obs = [case['first']] + [s['observation'] for s in steps]
  for (_, start, length, _), exp in zip(w.items, case['expected']):
    R, D = onstep.nstep_return([np.float32(x[2]) for x in w.steps[start:start + length]],
                               [np.float32(x[3]) for x in w.steps[start:start + length]],
                               case['additional_discount'])
    assert w.steps[start][0] == exp[0] == obs[start] and w.steps[start + length - 1][4] == exp[4]
    np.testing.assert_array_almost_equal([R, D], [exp[2], exp[3]])
    if len(exp) > 5:
      assert w.steps[start][5] == exp[5]
  adder.add_first(dm_env.restart(case['first']))
  adder.add(0, ts(steps[0]), extras=steps[0]['extras'] or ())
  assert len(client.writers) == 2
  assert client.writers[1].closed == (steps[0]['kind'] == 'term')


def test_protocol_errors_match_reference():
  adder = adders.NStepTransitionAdder(FakeClient(), 3, 0.99)
  with pytest.raises(ValueError):
    adder.add(0, dm_env.transition(0., 1))          # base.py:154-155
  with pytest.raises(ValueError):
    adder.add_first(dm_env.transition(0., 1))       # base.py:136-138
  adder.add_first(dm_env.restart(0))
  with pytest.raises(ValueError):
    adder.add_first(dm_env.restart(0))              # base.py:140-143
  with pytest.raises(ValueError):
    adders.NStepTransitionAdder(FakeClient(), 0, 0.99)


def test_custom_priority_fn_sees_stacked_window():
  seen = []

  def fn(x):
    seen.append(x)
    return float(np.abs(x.rewards).sum())

  client = FakeClient()
  adder = adders.NStepTransitionAdder(client, 2, 0.9, priority_fns={'t': fn})
  adder.add_first(dm_env.restart(np.zeros(3)))
  adder.add(1, dm_env.transition(2.0, np.ones(3)))
  adder.add(0, dm_env.termination(3.0, np.ones(3) * 2))
  assert seen[0].observations.shape == (2, 3) and seen[1].observations.shape == (3, 3)
  assert [p for _, _, _, p in client.writers[0].items] == [2.0, 5.0, 3.0]


def test_signature_order():
  from acme_b200 import specs
  spec = specs.EnvironmentSpec(specs.Array((2,), np.float32), specs.DiscreteArray(3), specs.Array((), np.float32),
                               specs.BoundedArray((), np.float32, 0., 1.))
  sig = adders.NStepTransitionAdder.signature(spec)
  assert sig == (spec.observations, spec.actions, spec.rewards, spec.discounts, spec.observations)
  sig = adders.NStepTransitionAdder.signature(spec, extras_spec={'s': specs.Array((), np.int32)})
  assert len(sig) == 6
