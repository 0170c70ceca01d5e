"""GPU parity of the replay path (K1 sample, K2 update, K3 gather + n-step) through the C ABI,
against oracle.replay / oracle.sumtree on the same inputs: indices, keys, probabilities, R and D
bit-exact; stored weights within 1 ulp of NumPy's pow (then synced, see helpers)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'nstep_cases.json')))


def _torch():
  import torch
  return torch


def gather_all(table, n_items):
  torch = _torch()
  from acme_b200 import replay
  ds = replay.ReplayDataset(table, n_items)
  ds.idx.copy_(torch.arange(n_items, dtype=torch.int64))
  table.flush()
  table.gather_into(ds.idx, ds.o_tm1, ds.a_tm1, ds.R, ds.D, ds.o_t)
  torch.cuda.synchronize()
  return ds.as_sample(table_size=n_items).data


@pytest.mark.parametrize('case', CASES, ids=[c['name'] for c in CASES])
def test_reference_golden_cases_through_the_gpu(case):
  """acme/adders/reverb/transition_test.py:29-170 with the CUDA ring + gather in the loop."""
  from acme_b200 import adders, dm_env, replay, specs
  first = case['first']
  is_dict = isinstance(first, dict)
  ospec = {'foo': specs.Array((), np.int64)} if is_dict else specs.Array((), np.int64)
  has_extras = case['steps'][0]['extras'] is not None
  espec = {'state': specs.Array((), np.int64)} if has_extras else ()
  spec = specs.EnvironmentSpec(ospec, specs.Array((), np.int64), specs.Array((), np.float64), specs.Array((), np.float64))
  table = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Prioritized(0.6), replay.selectors.Fifo(), 64,
                       replay.rate_limiters.MinSize(1), signature=adders.NStepTransitionAdder.signature(spec, espec),
                       max_window=case['n_step'], discount=case['additional_discount'])
  server = replay.Server([table])
  adder = adders.NStepTransitionAdder(replay.Client(server), case['n_step'], case['additional_discount'])
  adder.add_first(dm_env.restart(first))
  for s in case['steps']:
    ts = (dm_env.transition(s['reward'], s['observation'], s['discount']) if s['kind'] == 'mid'
          else dm_env.termination(s['reward'], s['observation']))
    adder.add(0, ts, extras=s['extras'] or ())
  exp = case['expected']
  assert table.size == len(exp)
  data = gather_all(table, len(exp))
  o0, a, R, D, o1 = data[:5]
  for i, e in enumerate(exp):
    got_o0 = {'foo': int(o0['foo'][i])} if is_dict else int(o0[i])
    got_o1 = {'foo': int(o1['foo'][i])} if is_dict else int(o1[i])
    assert got_o0 == e[0] and got_o1 == e[4] and int(a[i]) == e[1]
    np.testing.assert_array_almost_equal([float(R[i]), float(D[i])], [e[2], e[3]])
    if has_extras:
      assert int(data[5]['state'][i]) == e[5]['state']
  server.stop()


@pytest.mark.parametrize('n_step,alpha,max_size', [(3, 0.6, 500), (1, 1.0, 64), (5, 0.0, 300), (3, 0.6, 37)])
def test_lockstep_with_oracle(n_step, alpha, max_size):
  torch = _torch()
  from acme_b200 import replay
  import helpers
  rng = np.random.default_rng(n_step * 100 + max_size)
  shape, dtype, A = (8, 8, 4), np.uint8, 4
  spec, table, server, adder, oracle = helpers.make_pair(shape, dtype, A, n_step, 0.99, alpha, max_size)
  for ep in range(40):
    helpers.feed_episode(rng, adder, oracle, int(rng.integers(1, 25)), n_step, shape, dtype, A, terminal=ep % 3 != 0)
  table.flush()
  info = table.info()
  assert (info['size'], info['head_key'], info['tail_key']) == (oracle.size, oracle.item_head, oracle.item_tail)
  helpers.sync_oracle_leaves(table, oracle)
  for l in range(oracle.tree.L + 1):
    np.testing.assert_array_equal(table.read_tree_level(l)[:oracle.tree.levels[l].shape[0]], oracle.tree.levels[l])
  from oracle import sumtree as ost
  for l in range(1, oracle.tree.L + 1):   # the prefix lines the sampler reads = sequential fp32 scan of the children
    want = ost.seq_scan(oracle.tree.levels[l].reshape(-1, 32)).reshape(-1)
    np.testing.assert_array_equal(table.read_tree_prefix(l)[:want.shape[0]], want)
  B = 128
  ds = replay.ReplayDataset(table, B)
  for stratified in (True, False):
    ds.stratified = stratified
    u = rng.random(B, dtype=np.float32)
    ds.sample_raw(torch.as_tensor(u).cuda())
    torch.cuda.synchronize()
    keys, pos, prob = oracle.sample(u, stratified)
    np.testing.assert_array_equal(ds.idx.cpu().numpy(), pos)
    np.testing.assert_array_equal(ds.keys.cpu().numpy().view(np.uint64), keys)
    np.testing.assert_array_equal(ds.prob.cpu().numpy(), prob)
    o0, a, R, D, o1 = oracle.gather(pos)
    sample = ds.as_sample()
    # SampleInfo as the reference's fake dataset pins it (acme/testing/fakes.py:251-259): u64 / f64 / i64 / f64
    assert (sample.info.key.dtype, sample.info.probability.dtype, sample.info.table_size.dtype,
            sample.info.priority.dtype) == (torch.uint64, torch.float64, torch.int64, torch.float64)
    leaves = oracle.tree.levels[oracle.tree.L][pos].astype(np.float64)
    if alpha > 0:   # priority = stored weight ^ (1/alpha); new items enter with priority 1
      np.testing.assert_allclose(sample.info.priority.cpu().numpy(), leaves**(1.0 / alpha), rtol=1e-5)
    else:
      np.testing.assert_array_equal(sample.info.priority.cpu().numpy(), np.ones(B))
    s = sample.data
    np.testing.assert_array_equal(s[0].cpu().numpy(), o0)
    np.testing.assert_array_equal(s[1].cpu().numpy(), a)
    np.testing.assert_array_equal(s[2].cpu().numpy().view(np.uint32), R.view(np.uint32))   # bit-exact fp32
    np.testing.assert_array_equal(s[3].cpu().numpy().view(np.uint32), D.view(np.uint32))
    np.testing.assert_array_equal(s[4].cpu().numpy(), o1)
  assert (ds.prob > 0).all()
  # priority write-back: duplicates (last wins), zero priorities, stale and future keys
  keys = ds.keys.cpu().numpy().view(np.uint64).copy()
  keys[5] = keys[4]
  keys[6] = np.uint64(oracle.item_head + 7)
  if oracle.item_tail > 0:
    keys[7] = np.uint64(oracle.item_tail - 1)
  pr = np.abs(rng.standard_normal(B)).astype(np.float32)
  pr[10] = 0.0
  replay.Client(server).update_priorities(table.name, keys, pr.astype(np.float64))
  oracle.update_priorities(keys, pr)
  torch.cuda.synchronize()
  helpers.sync_oracle_leaves(table, oracle)
  for l in range(oracle.tree.L + 1):
    np.testing.assert_array_equal(table.read_tree_level(l)[:oracle.tree.levels[l].shape[0]], oracle.tree.levels[l])
  u = rng.random(B, dtype=np.float32)
  ds.sample_raw(torch.as_tensor(u).cuda())
  torch.cuda.synchronize()
  _, pos, prob = oracle.sample(u, False)
  np.testing.assert_array_equal(ds.idx.cpu().numpy(), pos)
  np.testing.assert_array_equal(ds.prob.cpu().numpy(), prob)
  server.stop()


def test_slot_ring_eviction_and_reset():
  """A slot ring smaller than 2*max_size evicts items whose first observation is overwritten."""
  torch = _torch()
  import helpers
  rng = np.random.default_rng(5)
  shape, dtype, A, n = (4, 4), np.float32, 3, 2
  spec, table, server, adder, oracle = helpers.make_pair(shape, dtype, A, n, 0.9, 0.6, max_size=200, slot_capacity=60,
                                                         stage_slots=16)
  for ep in range(30):
    helpers.feed_episode(rng, adder, oracle, int(rng.integers(1, 9)), n, shape, dtype, A)
  table.flush()
  info = table.info()
  assert (info['size'], info['head_key'], info['tail_key']) == (oracle.size, oracle.item_head, oracle.item_tail)
  assert oracle.size < 200 and oracle.item_tail > 0
  helpers.sync_oracle_leaves(table, oracle)
  from acme_b200 import replay
  ds = replay.ReplayDataset(table, 64)
  u = rng.random(64, dtype=np.float32)
  ds.sample_raw(torch.as_tensor(u).cuda())
  torch.cuda.synchronize()
  keys, pos, prob = oracle.sample(u, True)
  np.testing.assert_array_equal(ds.idx.cpu().numpy(), pos)
  np.testing.assert_array_equal(ds.keys.cpu().numpy().view(np.uint64), keys)
  o0, a, R, D, o1 = oracle.gather(pos)
  s = ds.as_sample().data
  np.testing.assert_array_equal(s[0].cpu().numpy(), o0)
  np.testing.assert_array_equal(s[4].cpu().numpy(), o1)
  np.testing.assert_array_equal(s[2].cpu().numpy(), R)
  replay.Client(server).reset(table.name)
  assert table.size == 0 and table.info()['total_mass'] == 0.0
  with pytest.raises(Exception):
    next(iter(ds))
  server.stop()


@pytest.mark.parametrize('N,B', [(1000, 256), (1 << 20, 256), (1 << 20, 65536), (3_000_000, 1 << 17)])
def test_priority_only_tree_sample_and_update(N, B):
  """BASELINE config 4 shape: priorities only, small and large batches (both K1/K2 code paths)."""
  torch = _torch()
  from acme_b200 import replay
  from oracle import sumtree
  rng = np.random.default_rng(N % 1000 + B)
  table = replay.Table.priorities_only('t', 0.6, N)
  w = np.abs(rng.standard_normal(N)).astype(np.float32)
  w[rng.integers(0, N, N // 10)] = 0.0
  table.set_weights(torch.as_tensor(w).cuda())
  tree = sumtree.SumTree(N)
  tree.levels[tree.L][:N] = w
  tree.rebuild()
  for l in range(tree.L + 1):
    np.testing.assert_array_equal(table.read_tree_level(l)[:tree.levels[l].shape[0]], tree.levels[l])
  dev = torch.device('cuda')
  idx = torch.empty(B, dtype=torch.int64, device=dev)
  keys = torch.empty(B, dtype=torch.uint64, device=dev)
  prob = torch.empty(B, dtype=torch.float32, device=dev)
  for stratified in (True, False):
    u = rng.random(B, dtype=np.float32)
    table.sample_into(torch.as_tensor(u).cuda(), idx, keys, prob, stratified)
    torch.cuda.synchronize()
    pos, p = tree.sample(u, stratified)
    np.testing.assert_array_equal(idx.cpu().numpy(), pos)
    np.testing.assert_array_equal(prob.cpu().numpy(), p)
    assert (w[pos] > 0).all()
  # update the sampled items with new priorities (alpha = 0.6 applied on the device)
  pr = np.abs(rng.standard_normal(B)).astype(np.float32)
  table.update_priorities_device(keys, torch.as_tensor(pr).cuda())
  torch.cuda.synchronize()
  k = keys.cpu().numpy().view(np.uint64).astype(np.int64)
  tree.set_leaves(k, sumtree.weight_from_priority(pr, 0.6))
  got = table.read_tree_level(tree.L)[:tree.levels[tree.L].shape[0]]
  np.testing.assert_allclose(got, tree.levels[tree.L], rtol=2.5e-7)
  tree.levels[tree.L][:] = got
  tree.rebuild()
  for l in range(tree.L + 1):
    np.testing.assert_array_equal(table.read_tree_level(l)[:tree.levels[l].shape[0]], tree.levels[l])
  table.close()


def test_empirical_frequencies_follow_priorities():
  """Closed-form property: P(i) = w_i / sum(w) (Prioritized selector contract)."""
  torch = _torch()
  from acme_b200 import replay
  N, B = 64, 1 << 18
  table = replay.Table.priorities_only('t', 1.0, N)
  w = np.arange(N, dtype=np.float32)
  table.set_weights(torch.as_tensor(w).cuda())
  dev = torch.device('cuda')
  idx = torch.empty(B, dtype=torch.int64, device=dev)
  prob = torch.empty(B, dtype=torch.float32, device=dev)
  u = torch.as_tensor(np.random.default_rng(0).random(B, dtype=np.float32)).cuda()
  table.sample_into(u, idx, None, prob, stratified=False)
  freq = np.bincount(idx.cpu().numpy(), minlength=N) / B
  np.testing.assert_allclose(freq, w / w.sum(), atol=2e-3)
  assert freq[0] == 0
  table.close()


def test_stream_append_matches_adder():
  """b200rl_writer_append_stream enumerates exactly the adder's items (bulk feeder path)."""
  torch = _torch()
  import ctypes as C
  from acme_b200 import _capi
  import helpers
  rng = np.random.default_rng(11)
  shape, dtype, A, n = (8, 8, 4), np.uint8, 4, 3
  spec, table, server, adder, oracle = helpers.make_pair(shape, dtype, A, n, 0.99, 0.6, max_size=400)
  obs, act, rew, disc, first, last = [], [], [], [], [], []
  for ep in range(12):
    T = int(rng.integers(1, 20))
    w = oracle.writer()
    o = helpers.random_obs(rng, shape, dtype)
    for k in range(T + 1):
      a, r, d = np.int32(rng.integers(A)), np.float32(rng.choice([-1., 0., 1.])), np.float32(0. if k == T - 1 else 1.)
      obs.append(o); act.append(a); rew.append(r); disc.append(d); first.append(k == 0); last.append(k == T)
      if k < T:
        o2 = helpers.random_obs(rng, shape, dtype)
        oracle.append(w, o, a, r, d, o2)
        oracle.create_item(w, min(k + 1, n), 1.0)
        if k == T - 1:
          m = min(n, T)
          for j in range(1, m):
            oracle.create_item(w, m - j, 1.0)
        o = o2
  obs = np.stack(obs); act = np.asarray(act, np.int32); rew = np.asarray(rew, np.float32); disc = np.asarray(disc, np.float32)
  first = np.asarray(first, np.uint8); last = np.asarray(last, np.uint8)
  wid = C.c_int32()
  _capi.call('b200rl_writer_open', table.handle, C.byref(wid))
  half = len(obs) // 2   # two calls, the second with device-resident observations, cut mid-episode
  _capi.call('b200rl_writer_append_stream', table.handle, wid.value, half, obs[:half].ctypes.data, 0, act[:half].ctypes.data,
             rew[:half].ctypes.data, disc[:half].ctypes.data, first[:half].ctypes.data, last[:half].ctypes.data, n, 1.0,
             _capi.current_stream())
  dobs = torch.as_tensor(obs[half:]).cuda()
  _capi.call('b200rl_writer_append_stream', table.handle, wid.value, len(obs) - half, dobs.data_ptr(), 1,
             act[half:].ctypes.data, rew[half:].ctypes.data, disc[half:].ctypes.data, first[half:].ctypes.data,
             last[half:].ctypes.data, n, 1.0, _capi.current_stream())
  torch.cuda.synchronize()
  assert table.size == oracle.size
  data = gather_all(table, oracle.size)
  o0, a, R, D, o1 = oracle.gather(np.arange(oracle.size))
  np.testing.assert_array_equal(data[0].cpu().numpy(), o0)
  np.testing.assert_array_equal(data[1].cpu().numpy(), a)
  np.testing.assert_array_equal(data[2].cpu().numpy(), R)
  np.testing.assert_array_equal(data[3].cpu().numpy(), D)
  np.testing.assert_array_equal(data[4].cpu().numpy(), o1)
  server.stop()


def test_philox_uniforms_are_reproducible():
  torch = _torch()
  from acme_b200 import _capi
  out = torch.empty(4096, dtype=torch.float32, device='cuda')
  step = torch.zeros(1, dtype=torch.int64, device='cuda')
  _capi.call('b200rl_uniform', out.data_ptr(), 4096, 1234, step.data_ptr(), 0, _capi.current_stream())
  a = out.cpu().numpy().copy()
  _capi.call('b200rl_uniform', out.data_ptr(), 4096, 1234, step.data_ptr(), 0, _capi.current_stream())
  np.testing.assert_array_equal(a, out.cpu().numpy())
  _capi.call('b200rl_uniform', out.data_ptr(), 4096, 1234, step.data_ptr(), 1, _capi.current_stream())
  b = out.cpu().numpy()
  assert (a != b).mean() > 0.99 and a.min() >= 0 and a.max() < 1 and abs(a.mean() - 0.5) < 0.03


def test_table_save_restore_resumes_exactly():
  """a table restored from save() continues exactly like the uninterrupted one -- same keys, FIFO position,
  priorities and n-step windows, including an episode that is open across the checkpoint."""
  torch = _torch()
  import helpers
  from acme_b200 import dm_env, replay
  rng = np.random.default_rng(12)
  shape, dtype, A, n = (8, 8, 4), np.uint8, 4, 3
  spec, t1, s1, adder1, _ = helpers.make_pair(shape, dtype, A, n, 0.99, 0.6, max_size=120)
  _, t2, s2, adder2, _ = helpers.make_pair(shape, dtype, A, n, 0.99, 0.6, max_size=120)

  class Null:   # the helper feeds an oracle as well; not needed here
    def writer(self): return None
    def append(self, *a): pass
    def create_item(self, *a): pass
    def close(self, *a): pass

  for ep in range(12):     # > max_size items: the FIFO has wrapped
    helpers.feed_episode(rng, adder1, Null(), int(rng.integers(3, 25)), n, shape, dtype, A)
  # an episode left open across the checkpoint
  obs = [helpers.random_obs(rng, shape, dtype) for _ in range(12)]
  adder1.add_first(dm_env.restart(obs[0]))
  for k in range(1, 5):
    adder1.add(np.int32(k % A), dm_env.transition(np.float32(k), obs[k], np.float32(1.)))
  keys = np.arange(t1.info()['tail_key'], t1.info()['head_key'], dtype=np.uint64)
  replay.Client(s1).update_priorities(t1.name, keys, np.abs(rng.standard_normal(keys.shape[0])))
  state = t1.save()
  t2.restore(state)
  assert t2.info() == t1.info()
  # the second adder continues the open episode: give it the same writer state by construction
  adder2._buffer = adder1._buffer.__class__(adder1._buffer, maxlen=adder1._buffer.maxlen)
  adder2._next_observation, adder2._start_of_episode = adder1._next_observation, adder1._start_of_episode
  w1, w2 = adder1._writer, adder2._writer            # the property creates adder2's writer (lazily, one per episode)
  w2._ids = dict(w1._ids)                            # same writer slot in the restored table: its window came with the blob
  w2._started = set(w1._started)
  for k in range(5, 11):
    for ad in (adder1, adder2):
      last = k == 10
      ts = dm_env.TimeStep(dm_env.StepType.LAST if last else dm_env.StepType.MID, np.float32(k), np.float32(0. if last else 1.), obs[k])
      ad.add(np.int32(k % A), ts)
  rng2 = np.random.default_rng(5)
  for ep in range(3):
    T = int(rng2.integers(3, 20))
    seeds = int(rng2.integers(1 << 30))
    for ad in (adder1, adder2):
      helpers.feed_episode(np.random.default_rng(seeds), ad, Null(), T, n, shape, dtype, A)
  t1.flush(); t2.flush()
  assert t2.info() == t1.info()
  L = t1.tree_levels()[0]
  for l in range(L + 1):
    np.testing.assert_array_equal(t2.read_tree_level(l), t1.read_tree_level(l))
  B = 64
  d1, d2 = replay.ReplayDataset(t1, B), replay.ReplayDataset(t2, B)
  u = torch.as_tensor(rng.random(B, dtype=np.float32)).cuda()
  d1.sample_raw(u); d2.sample_raw(u)
  torch.cuda.synchronize()
  for a, b in ((d1.idx, d2.idx), (d1.prob, d2.prob), (d1.o_both, d2.o_both), (d1.a_tm1, d2.a_tm1), (d1.R, d2.R), (d1.D, d2.D)):
    assert torch.equal(a, b)
  assert torch.equal(d1.keys.view(torch.int64), d2.keys.view(torch.int64))
  s1.stop(); s2.stop()


@pytest.mark.parametrize('hw,n_step,max_size', [(84, 3, 400), (20, 5, 150), (8, 1, 64)])
def test_frame_dedup_ring_rebuilds_stacks(hw, n_step, max_size):
  """a table with frame_stack=4 stores one frame per step; K3 must hand back exactly the stacks the
  FrameStacker (frame_stacking.py:64-88) produced -- including the blank frames at episode starts, items whose two
  observations straddle short episodes, and after the item FIFO has wrapped.  Checked against the oracle table, which
  stores every stack whole, for both gather forms (plain, and with conv1's bf16 row image where the geometry has one)."""
  import ctypes
  torch = _torch()
  import helpers
  from acme_b200 import _capi, replay
  rng = np.random.default_rng(hw + n_step)
  shape, A = (hw, hw, 4), 4
  spec, table, server, adder, oracle = helpers.make_pair(shape, np.uint8, A, n_step, 0.99, 0.6, max_size, frame_stack=4)
  for ep in range(30):
    helpers.feed_stacked_episode(rng, adder, oracle, int(rng.integers(1, 40)), n_step, (hw, hw), 4, A, terminal=ep % 3 != 0)
  table.flush()
  info = table.info()
  assert (info['size'], info['head_key'], info['tail_key']) == (oracle.size, oracle.item_head, oracle.item_tail)
  helpers.sync_oracle_leaves(table, oracle)
  B = 128
  ds = replay.ReplayDataset(table, B)
  for trial in range(3):
    u = rng.random(B, dtype=np.float32)
    ds.sample_raw(torch.as_tensor(u).cuda())
    torch.cuda.synchronize()
    keys, pos, prob = oracle.sample(u, True)
    np.testing.assert_array_equal(ds.idx.cpu().numpy(), pos)
    o0, a, R, D, o1 = oracle.gather(pos)
    s = ds.as_sample().data
    np.testing.assert_array_equal(s[0].cpu().numpy(), o0)          # stacks rebuilt from single frames: bit-exact
    np.testing.assert_array_equal(s[4].cpu().numpy(), o1)
    np.testing.assert_array_equal(s[1].cpu().numpy(), a)
    np.testing.assert_array_equal(s[2].cpu().numpy().view(np.uint32), R.view(np.uint32))
    np.testing.assert_array_equal(s[3].cpu().numpy().view(np.uint32), D.view(np.uint32))
  assert (o0[..., 0] == 0).all(axis=(1, 2)).any(), 'no episode-start stack in the sample: the blank-frame case was not exercised'
  if hw in (84, 20):        # geometries with a conv1 row image: the fused form writes the same image as gather + conversion
    from acme_b200 import networks
    oh, pad = networks.tf_same_pad(hw, 8, 4)
    g = lambda b: _capi.ConvGeom(B=b, H=hw, W=hw, C=4, kh=8, kw=8, stride=4, pad_top=pad, pad_left=pad, OH=oh, OW=oh, Cout=32)
    fb = int(_capi.load().b200rl_conv2d_rows_bf16_bytes(ctypes.byref(g(1))))
    rows = torch.zeros(2 * B * fb, dtype=torch.uint8, device='cuda')
    both = ds.o_both.clone()
    ds.o_both.zero_()
    ds.gather_only((rows.data_ptr(), rows.data_ptr() + B * fb, g(1)))
    want = torch.zeros(2 * B * fb, dtype=torch.uint8, device='cuda')
    _capi.call('b200rl_conv2d_rows_bf16_from_u8', both.data_ptr(), g(2 * B), want.data_ptr(), want.numel(), _capi.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(ds.o_both, both) and torch.equal(rows, want)
  # the ring really is four times smaller
  seg = table._segment(0)
  assert seg.numel() == table.slot_capacity * (-(-(hw * hw) // 16) * 16)
  server.stop()


def test_concurrent_actor_threads_feed_one_table():
  """SURVEY §8f-4 (async actor feed): several actor threads, each with its own adder / writer, insert into ONE table
  while the learner thread flushes and samples -- Reverb's server is thread-safe and Acme runs actors against it
  concurrently; here every C entry point takes the shard's lock.  Observations encode (actor, episode, t), so every
  gathered n-step window can be checked: same actor, same episode, o_t exactly `len` steps after o_tm1, R / D the n-step
  values of that stretch."""
  import threading
  torch = _torch()
  import helpers
  from acme_b200 import adders, dm_env, replay
  n_step, A, actors, episodes, T = 3, 4, 3, 25, 17
  spec, table, server, adder0, _ = helpers.make_pair((4,), np.float32, A, n_step, 0.5, 0.6, max_size=5000)
  errors = []

  def actor(aid):
    try:
      ad = adders.NStepTransitionAdder(replay.Client(server), n_step=n_step, discount=0.5)
      for ep in range(episodes):
        ad.add_first(dm_env.restart(np.array([aid, ep, 0, 0], np.float32)))
        for t in range(1, T + 1):
          last = t == T
          ts = dm_env.TimeStep(dm_env.StepType.LAST if last else dm_env.StepType.MID, np.float32(t), np.float32(0. if last else 1.),
                               np.array([aid, ep, t, 0], np.float32))
          ad.add(np.int32(t % A), ts)
    except Exception as e:   # pragma: no cover
      errors.append(e)

  threads = [threading.Thread(target=actor, args=(i,)) for i in range(actors)]
  for th in threads:
    th.start()
  B = 64
  ds = None
  seen = 0
  while any(th.is_alive() for th in threads) or seen < 20:
    table.flush()
    if table.size >= 1:
      ds = ds or replay.ReplayDataset(table, B, seed=3)
      ds.sample_raw()
      torch.cuda.synchronize()
      o0 = ds.o_tm1.view(torch.float32).view(B, 4).cpu().numpy()
      o1 = ds.o_t.view(torch.float32).view(B, 4).cpu().numpy()
      R, D = ds.R.cpu().numpy(), ds.D.cpu().numpy()
      np.testing.assert_array_equal(o0[:, :2], o1[:, :2])                       # same actor, same episode
      length = (o1[:, 2] - o0[:, 2]).astype(int)
      assert ((length >= 1) & (length <= n_step)).all()
      for b in range(B):   # n-step return of rewards t0+1 .. t0+len with discount 0.5 and env discounts 1 (0 at the end)
        t0, ln = int(o0[b, 2]), int(length[b])
        want_R = sum((0.5**j) * (t0 + 1 + j) for j in range(ln))
        want_D = (0.5**(ln - 1)) * (0. if t0 + ln == T else 1.)
        assert abs(R[b] - want_R) < 1e-5 and abs(D[b] - want_D) < 1e-7, (b, t0, ln, R[b], want_R, D[b], want_D)
      seen += 1
  for th in threads:
    th.join()
  assert not errors, errors
  table.flush()
  assert table.size == actors * episodes * (T + n_step - 1)
  server.stop()
