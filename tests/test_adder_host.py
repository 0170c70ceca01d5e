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


# ---- randomised: product adder (windows over appended steps) == the reference adder restatement (materialised items)
def _fuzz_once(rng):
  n_step = int(rng.integers(1, 7))
  gamma = float(rng.choice([1.0, 0.99, 0.5]))
  product_client, reference_client = FakeClient(), onstep.RecordingClient()
  product = adders.NStepTransitionAdder(product_client, n_step, gamma)
  reference = onstep.ReferenceAdder(reference_client, n_step, gamma)
  for _ in range(int(rng.integers(1, 4))):                       # a few episodes through the same adders
    T = int(rng.integers(1, 15))
    first = dm_env.restart(float(rng.standard_normal()))
    product.add_first(first)
    reference.add_first(first)
    for t in range(T):
      obs, rew = float(rng.standard_normal()), np.float32(rng.standard_normal())
      if t < T - 1:
        step = dm_env.transition(reward=rew, observation=obs, discount=np.float32(rng.choice([1.0, 0.9, 0.0])))
      elif rng.random() < 0.5:
        step = dm_env.termination(reward=rew, observation=obs)
      else:
        step = dm_env.truncation(reward=rew, observation=obs, discount=np.float32(1.0))
      a = int(rng.integers(0, 5))
      product.add(a, step)
      reference.add(a, step)
  assert len(product_client.writers) == len(reference_client.writers)
  for pw, rw in zip(product_client.writers, reference_client.writers):
    assert pw.closed and rw.closed
    assert len(pw.items) == len(rw.priorities)
    for (table, start, length, prio), (rtable, item, rprio) in zip(pw.items, rw.priorities):
      window = pw.steps[start:start + length]
      R, D = onstep.nstep_return([np.float32(s[2]) for s in window], [np.float32(s[3]) for s in window], np.float32(gamma))
      assert (table, prio) == (rtable, rprio)
      assert window[0][0] == item[0] and window[0][1] == item[1] and window[-1][4] == item[4]
      assert np.float32(R) == np.float32(item[2]) and np.float32(D) == np.float32(item[3])   # bit-exact bookkeeping


def test_random_episodes_match_the_reference_adder():
  rng = np.random.default_rng(2024)
  for _ in range(300):
    _fuzz_once(rng)
