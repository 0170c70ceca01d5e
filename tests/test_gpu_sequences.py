"""GPU parity of sequence replay through the C ABI: SequenceAdder -> HBM step ring -> sequence items ->
K3 sequence gather, against the reference's golden cases (`acme/adders/reverb/sequence_test.py:25-181`) and against
oracle.sequence on random episodes (bit-exact: every byte of every step of every item); the recurrent learner's priority
mix and importance weights (`acme/agents/tf/r2d2/learning.py:170-176,230-236`) against their restatements."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'sequence_cases.json')))


def _table(spec, extras_spec, L, max_size=256, slot_capacity=None, alpha=0.6):
  from acme_b200 import adders, replay
  table = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Prioritized(alpha), replay.selectors.Fifo(), max_size,
                       replay.rate_limiters.MinSize(1), signature=adders.SequenceAdder.signature(spec, extras_spec),
                       max_window=L, slot_capacity=slot_capacity)
  return table, replay.Server([table])


def gather_items(table, positions, L, time_major=False):
  import torch
  from acme_b200 import replay
  ds = replay.ReplayDataset(table, len(positions), sequence_length=L, time_major=time_major)
  ds.idx.copy_(torch.as_tensor(np.asarray(positions, np.int64)))
  ds.keys.zero_()
  ds.prob.fill_(1.0)
  table.flush()
  ds.gather_only()
  torch.cuda.synchronize()
  return ds.as_sample(table_size=len(positions)).data


@pytest.mark.parametrize('case', CASES, ids=[c['name'] for c in CASES])
def test_reference_golden_cases_through_the_gpu(case):
  from acme_b200 import adders, dm_env, replay, specs
  L = case['sequence_length']
  spec = specs.EnvironmentSpec(specs.Array((), np.int64), specs.Array((), np.int64), specs.Array((), np.float64),
                               specs.Array((), np.float64))
  table, server = _table(spec, (), L)
  adder = adders.SequenceAdder(replay.Client(server), sequence_length=L, period=case['period'],
                               pad_end_of_episode=case['pad_end_of_episode'])
  adder.add_first(dm_env.restart(case['first']))
  for s in case['steps']:
    ts = (dm_env.transition(s['reward'], s['observation'], s['discount']) if s['kind'] == 'mid'
          else dm_env.termination(s['reward'], s['observation']))
    adder.add(s['action'], ts)
  # the reference code writes no item in 'EarlyTerminationNoPadding' (see tests/test_sequence_host.py)
  exp = [] if case['name'] == 'EarlyTerminationNoPadding' else case['expected']
  assert table.size == len(exp)
  if exp:
    d = gather_items(table, range(len(exp)), L)
    for i, seq in enumerate(exp):
      got = [[int(d.observation[i, t]), int(d.action[i, t]), float(d.reward[i, t]), float(d.discount[i, t]),
              bool(d.start_of_episode[i, t]), []] for t in range(L)]
      assert got == seq
  server.stop()


def _feed(rng, adder, oracle, T, obs_shape, act_dim, terminal=True):
  from acme_b200 import dm_env
  o = rng.integers(0, 256, obs_shape, dtype=np.uint8)
  adder.add_first(dm_env.restart(o))
  oracle.add_first(dm_env.restart(o))
  for k in range(1, T + 1):
    a = rng.uniform(-1, 1, act_dim).astype(np.float32)
    ex = {'core': rng.standard_normal(3).astype(np.float32), 'id': np.int32(rng.integers(1 << 20))}
    r = np.float32(rng.choice([-1., 0., 1., 0.5, 2.5]))
    last = k == T
    d = np.float32(0. if (last and terminal) else rng.choice([1., 1., 0.9]))
    o = rng.integers(0, 256, obs_shape, dtype=np.uint8)
    ts = dm_env.TimeStep(dm_env.StepType.LAST if last else dm_env.StepType.MID, r, d, o)
    adder.add(a, ts, extras=ex)
    oracle.add(a, ts, extras=ex)


def _compare(d, i, t, step, time_major):
  ix = (t, i) if time_major else (i, t)
  assert np.array_equal(d.observation[ix].cpu().numpy(), np.asarray(step.observation))
  assert np.array_equal(d.action[ix].cpu().numpy(), np.asarray(step.action, np.float32))
  assert float(d.reward[ix]) == float(np.float32(step.reward)) and float(d.discount[ix]) == float(np.float32(step.discount))
  assert bool(d.start_of_episode[ix]) == bool(step.start_of_episode)
  assert np.array_equal(d.extras['core'][ix].cpu().numpy(), np.asarray(step.extras['core'], np.float32))
  assert int(d.extras['id'][ix]) == int(step.extras['id'])


@pytest.mark.parametrize('L,period,pad,time_major', [(4, 1, True, False), (5, 3, True, True), (6, 6, False, False), (3, 2, True, True)])
def test_random_episodes_match_oracle(L, period, pad, time_major):
  """Random episodes (some shorter than a sequence: padding; some long: overlapping windows) through the product adder +
  ring + sequence gather == the oracle adder's materialised sequences, step by step, byte by byte."""
  from acme_b200 import adders, replay, specs
  from oracle import nstep as onstep
  from oracle import sequence as oseq
  rng = np.random.default_rng(L * 100 + period)
  obs_shape, act_dim = (6, 6, 2), 3
  spec = specs.EnvironmentSpec(specs.Array(obs_shape, np.uint8), specs.BoundedArray((act_dim,), np.float32, -1., 1.),
                               specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
  espec = {'core': specs.Array((3,), np.float32), 'id': specs.Array((), np.int32)}
  table, server = _table(spec, espec, L, max_size=512)
  adder = adders.SequenceAdder(replay.Client(server), L, period, pad_end_of_episode=pad)
  client = onstep.RecordingClient()
  oracle = oseq.ReferenceSequenceAdder(client, L, period, pad)
  for T in (1, 2, L - 1, L, L + 1, 3 * L + 2, 11, 2):
    _feed(rng, adder, oracle, max(T, 1), obs_shape, act_dim, terminal=bool(rng.integers(2)))
  items = [it if isinstance(it, list) else [it] for w in client.writers for _, it, _ in w.priorities]
  assert table.size == len(items) >= 6
  d = gather_items(table, range(len(items)), L, time_major)
  assert d.reward.shape == ((L, len(items)) if time_major else (len(items), L))
  for i, seq in enumerate(items):
    assert len(seq) == L
    for t, step in enumerate(seq):
      _compare(d, i, t, step, time_major)
  server.stop()


def test_interleaved_writers_and_ring_wrap():
  """Two actors feeding one table alternately: their steps interleave in the slot ring, so an item's slots are not
  consecutive and the gather follows the per-slot links; a small ring makes windows wrap around its end and evicts the
  oldest items.  Live items still match the oracle's."""
  from acme_b200 import adders, dm_env, replay, specs
  from oracle import nstep as onstep
  from oracle import sequence as oseq
  rng = np.random.default_rng(5)
  L, period, obs_shape, act_dim = 4, 2, (4, 4, 1), 2
  spec = specs.EnvironmentSpec(specs.Array(obs_shape, np.uint8), specs.BoundedArray((act_dim,), np.float32, -1., 1.),
                               specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
  espec = {'core': specs.Array((3,), np.float32), 'id': specs.Array((), np.int32)}
  table, server = _table(spec, espec, L, max_size=64, slot_capacity=61)
  cl = replay.Client(server)
  prod = [adders.SequenceAdder(cl, L, period), adders.SequenceAdder(cl, L, period)]
  rec = onstep.RecordingClient()
  orac = [oseq.ReferenceSequenceAdder(rec, L, period), oseq.ReferenceSequenceAdder(rec, L, period)]
  left = [0, 0]
  for _ in range(150):
    k = int(rng.integers(2))
    if left[k] == 0:
      o = rng.integers(0, 256, obs_shape, dtype=np.uint8)
      prod[k].add_first(dm_env.restart(o))
      orac[k].add_first(dm_env.restart(o))
      left[k] = int(rng.integers(1, 12))
    a = rng.uniform(-1, 1, act_dim).astype(np.float32)
    ex = {'core': rng.standard_normal(3).astype(np.float32), 'id': np.int32(rng.integers(1, 1 << 30))}
    left[k] -= 1
    last = left[k] == 0
    o = rng.integers(0, 256, obs_shape, dtype=np.uint8)
    ts = dm_env.TimeStep(dm_env.StepType.LAST if last else dm_env.StepType.MID, np.float32(rng.integers(3)), np.float32(not last), o)
    prod[k].add(a, ts, extras=ex)
    orac[k].add(a, ts, extras=ex)
  # an item is identified by the (random, non-zero) ids of its first two steps
  lookup = {(int(it[0].extras['id']), int(it[1].extras['id'])): it for w in rec.writers for _, it, _ in w.priorities}
  info = table.info()
  n_live = info['size']
  assert 0 < n_live < sum(len(w.priorities) for w in rec.writers)     # the small ring evicted the oldest items
  tail, head = info['tail_key'], info['head_key']
  assert head - tail == n_live
  pos = [(k % table.max_size) for k in range(tail, head)]
  d = gather_items(table, pos, L)
  for i in range(n_live):
    key = (int(d.extras['id'][i, 0]), int(d.extras['id'][i, 1]))
    assert key in lookup, 'gathered a sequence the oracle never produced'
    for t, step in enumerate(lookup[key]):
      _compare(d, i, t, step, False)
  server.stop()


def test_sampled_sequences_and_priority_update():
  """make_reverb_dataset(sequence_length=...) end to end: K1 sample -> K3 sequence gather -> info tiled over time;
  compute_priority / importance weights kernels == oracle; TFClient.update_priorities(keys[:, 0], ...) lands."""
  import torch
  from acme_b200 import _capi, adders, replay, specs
  from oracle import nstep as onstep
  from oracle import sequence as oseq
  rng = np.random.default_rng(9)
  L, period, B = 8, 4, 32
  obs_shape, act_dim = (10,), 2
  spec = specs.EnvironmentSpec(specs.Array(obs_shape, np.uint8), specs.BoundedArray((act_dim,), np.float32, -1., 1.),
                               specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
  espec = {'core': specs.Array((3,), np.float32), 'id': specs.Array((), np.int32)}
  table, server = _table(spec, espec, L, max_size=128)
  client = replay.Client(server)
  adder = adders.SequenceAdder(client, L, period)
  oracle = oseq.ReferenceSequenceAdder(onstep.RecordingClient(), L, period)
  for T in (30, 17, 9, 40):
    _feed(rng, adder, oracle, T, obs_shape, act_dim)
  ds = replay.make_reverb_dataset(client=client, batch_size=B, sequence_length=L, seed=3)
  sample = next(ds)
  torch.cuda.synchronize()
  assert sample.info.key.shape == (B, L) and sample.info.probability.dtype == torch.float64
  assert sample.data.observation.shape == (B, L) + obs_shape and sample.data.extras['core'].shape == (B, L, 3)
  keys = sample.info.key[:, 0]                                             # r2d2/learning.py:223
  assert torch.equal(sample.info.key, keys.unsqueeze(1).expand(-1, L))
  # learner arithmetic
  err = torch.as_tensor(rng.standard_normal((L, B)).astype(np.float32)).cuda()
  prio = torch.empty(B, device='cuda')
  _capi.call('b200rl_seq_priority', L, B, _capi.ptr(err), 0.9, _capi.ptr(prio), _capi.current_stream())
  np.testing.assert_array_equal(prio.cpu().numpy(), oseq.compute_priority(err.cpu().numpy(), 0.9))
  w = torch.empty(B, device='cuda')
  _capi.call('b200rl_seq_is_weights', B, _capi.ptr(ds.prob), float(table.max_size), 0.6, _capi.ptr(w), _capi.current_stream())
  np.testing.assert_allclose(w.cpu().numpy(), oseq.importance_weights(ds.prob.cpu().numpy(), table.max_size, 0.6), rtol=1e-6)
  assert float(w.max()) == 1.0
  # priority write-back, then every later draw must see the new weights: make one item dominate
  big = torch.full((B,), 1e-6, device='cuda')
  big[0] = 1e6
  client.update_priorities(replay.DEFAULT_PRIORITY_TABLE, keys, big)
  s2 = next(ds)
  torch.cuda.synchronize()
  hit = (s2.info.key[:, 0] == keys[0]).float().mean().item()
  assert hit > 0.9
  server.stop()
