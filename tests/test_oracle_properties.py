"""Closed-form and cross-implementation checks of the oracle pieces that have no golden vectors in the reference tree
(Reverb's selector, trfl, Sonnet's Adam: "parity unpinned" in DESIGN.md §5).  CPU only.  They pin the oracle to
independent statements of the same maths: brute-force formulas, PyTorch's own optimizer / conv, distribution identities."""
import numpy as np
import pytest

from oracle import learner as olearner
from oracle import losses
from oracle import nets as onets
from oracle import sumtree as ost

f32 = np.float32


# ------------------------------------------------------------------ sum tree (contract of acme/agents/tf/dqn/agent.py:95-101)
def test_seq_scan_is_float32_cumsum():
  x = np.random.default_rng(0).random((7, 32)).astype(f32)
  np.testing.assert_array_equal(ost.seq_scan(x), np.cumsum(x, axis=-1, dtype=f32))


@pytest.mark.parametrize('capacity', [1, 31, 32, 33, 1000, 40_000])
def test_incremental_update_equals_rebuild(capacity):
  rng = np.random.default_rng(capacity)
  a, b = ost.SumTree(capacity), ost.SumTree(capacity)
  w = rng.random(capacity).astype(f32)
  a.set_leaves(np.arange(capacity), w)
  b.leaves[:capacity] = w
  b.rebuild()
  for la, lb in zip(a.levels, b.levels):
    np.testing.assert_array_equal(la, lb)
  # a second, sparse update with duplicates: last occurrence wins, untouched nodes keep their bits
  pos = rng.integers(0, capacity, 17)
  val = rng.random(17).astype(f32)
  a.set_leaves(pos, val)
  for p, v in zip(pos, val):
    b.leaves[p] = v
  b.rebuild()
  for la, lb in zip(a.levels, b.levels):
    np.testing.assert_array_equal(la, lb)


def test_sampling_follows_the_weights_and_skips_zeros():
  rng = np.random.default_rng(3)
  n = 200
  w = rng.random(n).astype(f32)
  w[rng.integers(0, n, 40)] = 0.
  t = ost.SumTree(n)
  t.set_leaves(np.arange(n), w)
  draws = 400_000
  idx, prob = t.sample(rng.random(draws).astype(f32), stratified=False)
  assert idx.max() < n and np.all(w[idx] > 0)                       # zero-weight items are never drawn
  np.testing.assert_allclose(prob, w[idx] / t.total, rtol=1e-6)     # reported probability = weight / mass
  freq = np.bincount(idx, minlength=n) / draws
  expect = w.astype(np.float64) / w.astype(np.float64).sum()
  assert np.abs(freq - expect).max() < 5 * np.sqrt(expect.max() / draws) + 1e-4


def test_stratified_draws_are_one_per_stratum_and_ordered():
  n, B = 5000, 64
  t = ost.SumTree(n)
  w = np.random.default_rng(4).random(n).astype(f32)
  t.set_leaves(np.arange(n), w)
  u = np.random.default_rng(5).random(B).astype(f32)
  idx, _ = t.sample(u, stratified=True)
  assert np.all(np.diff(idx) >= 0)                                  # targets increase with the stratum
  cdf = np.cumsum(w.astype(np.float64))
  lo = cdf[idx] - w[idx]
  strata = (np.arange(B) / B) * cdf[-1]
  assert np.all(lo <= strata + cdf[-1] / B + 1e-3 * cdf[-1] / B * B)  # each draw lies in (or on the edge of) its stratum


def test_same_index_as_a_binary_f64_tree_when_sums_are_exact():
  """With small integer weights every partial sum is exact in fp32 and f64, so the fan-out-32 fp32 tree and a Reverb-like
  binary f64 tree must pick the same item for the same target."""
  rng = np.random.default_rng(6)
  n = 4096
  w = rng.integers(0, 8, n).astype(f32)
  t = ost.SumTree(n)
  t.set_leaves(np.arange(n), w)
  b = ost.BinarySumTreeF64(n)
  b.build(w.astype(np.float64))
  # targets on a grid of exactly representable values strictly inside the mass
  mass = float(t.total)
  targets = (np.arange(1, 2000) * (mass / 2048.)).astype(f32)
  u = (targets / f32(mass)).astype(f32)
  assert np.all((u * f32(mass)).astype(f32) == targets)             # the grid survives the u * mass round trip
  i32, _ = t.sample(u, stratified=False)
  i64, _ = b.sample(u.astype(np.float64))
  np.testing.assert_array_equal(i32, i64)


# ------------------------------------------------------------------ losses
def test_huber_matches_its_definition():
  x = np.linspace(-3, 3, 601).astype(f32)
  for delta in (0.5, 1.0, 2.0):
    ref = np.where(np.abs(x) <= delta, 0.5 * x.astype(np.float64)**2, delta * (np.abs(x) - 0.5 * delta))
    np.testing.assert_allclose(losses.huber(x, delta), ref, rtol=1e-6, atol=1e-7)     # acme/tf/losses/huber.py:48-57


def test_double_q_td_against_a_loop():
  rng = np.random.default_rng(7)
  B, A = 33, 6
  q, qv, qs = (rng.standard_normal((B, A)).astype(f32) for _ in range(3))
  a = rng.integers(0, A, B)
  R, D = rng.standard_normal(B).astype(f32) * 2, rng.random(B).astype(f32)
  prob = rng.random(B) * 0.01 + 1e-4
  out = losses.dqn_loss(q, qv, qs, a, R, D, prob, gamma=0.99, huber_delta=1.0, is_exponent=0.2, max_abs_reward=1.0)
  for i in range(B):
    r = min(max(float(R[i]), -1.0), 1.0)
    target = r + float(D[i]) * 0.99 * float(qv[i, int(np.argmax(qs[i]))])   # select with the online net, evaluate with the target net
    assert abs(float(out['td'][i]) - (target - float(q[i, a[i]]))) < 1e-5
  w = (1.0 / prob)**0.2
  np.testing.assert_allclose(out['weight'], w / w.max(), rtol=1e-6)           # dqn/learning.py:138-140
  np.testing.assert_allclose(out['priority'], np.abs(out['td']), rtol=0)      # dqn/learning.py:151
  # gradient of mean(w * huber(td)) w.r.t. q_tm1 by finite differences
  eps = 1e-3
  for i in (0, 5, 17):
    qp, qm = q.copy(), q.copy()
    qp[i, a[i]] += eps
    qm[i, a[i]] -= eps
    lp = losses.dqn_loss(qp, qv, qs, a, R, D, prob, 0.99)['per_sample'].astype(np.float64).sum() / B
    lm = losses.dqn_loss(qm, qv, qs, a, R, D, prob, 0.99)['per_sample'].astype(np.float64).sum() / B
    assert abs((lp - lm) / (2 * eps) - float(out['dq_tm1'][i, a[i]])) < 2e-4


def test_c51_projection_identities():
  rng = np.random.default_rng(8)
  K, B = 51, 9
  z = np.linspace(-10, 10, K).astype(f32)
  p = rng.random((B, K)).astype(f32)
  p /= p.sum(axis=1, keepdims=True)
  # identity when the source support is the target support
  np.testing.assert_allclose(losses.l2_project(np.tile(z, (B, 1)), p, z), p, atol=1e-6)
  # mass is conserved, and so is the mean while nothing is clipped (linear interpolation between neighbours)
  zp = (0.3 + 0.9 * z[None, :] * np.ones((B, 1))).astype(f32)
  proj = losses.l2_project(zp, p, z)
  np.testing.assert_allclose(proj.sum(axis=1), 1.0, atol=1e-5)
  np.testing.assert_allclose((proj * z).sum(axis=1), (p * zp).sum(axis=1), atol=2e-4)
  # everything beyond the support piles up on the edge atoms
  far = losses.l2_project(np.full((1, K), 1e3, f32), p[:1], z)
  np.testing.assert_allclose(far[0, -1], 1.0, atol=1e-6)
  # brute force (Bellemare et al. 2017, Alg. 1) on a random case
  Dg, R = f32(0.97), rng.standard_normal(B).astype(f32)
  tz = np.clip(R[:, None] + Dg * z[None, :], z[0], z[-1]).astype(np.float64)
  dz = float(z[1] - z[0])
  brute = np.zeros((B, K))
  for b in range(B):
    for j in range(K):
      pos = min(max((tz[b, j] - float(z[0])) / dz, 0.0), K - 1.0)
      lo, hi = int(np.floor(pos)), int(np.ceil(pos))
      if lo == hi:
        brute[b, lo] += p[b, j]
      else:
        brute[b, lo] += p[b, j] * (hi - pos)
        brute[b, hi] += p[b, j] * (pos - lo)
  np.testing.assert_allclose(losses.l2_project((R[:, None] + Dg * z[None, :]).astype(f32), p, z), brute, atol=3e-5)


def test_categorical_loss_gradient_by_finite_differences():
  rng = np.random.default_rng(9)
  B, K = 4, 11
  z = np.linspace(-2, 2, K).astype(f32)
  lt, ln = rng.standard_normal((B, K)).astype(f32), rng.standard_normal((B, K)).astype(f32)
  R, Dg = rng.standard_normal(B).astype(f32) * 0.3, np.full(B, 0.9, f32)
  out = losses.categorical(lt, ln, z, R, Dg)
  eps = 1e-2
  for (b, k) in ((0, 0), (2, 5), (3, 10)):
    lp, lm = lt.copy(), lt.copy()
    lp[b, k] += eps
    lm[b, k] -= eps
    fd = (losses.categorical(lp, ln, z, R, Dg)['loss'].astype(np.float64).mean() -
          losses.categorical(lm, ln, z, R, Dg)['loss'].astype(np.float64).mean()) / (2 * eps)
    assert abs(fd - float(out['dlogits_tm1'][b, k])) < 2e-4


def test_dpg_and_global_norm_clipping():
  g = np.array([[3., 4.], [0.3, 0.4]], f32)
  da, clipped = losses.dpg_action_grad(g, clip=1.0, clip_norm=True)
  np.testing.assert_allclose(clipped, [[0.6, 0.8], [0.3, 0.4]], rtol=1e-6)      # tf.clip_by_norm (dpg.py:41-57)
  np.testing.assert_allclose(da, -clipped / 2, rtol=1e-6)
  out, norm = losses.clip_by_global_norm([np.array([3.], f32), np.array([4.], f32)], 2.5)
  assert abs(float(norm) - 5.0) < 1e-6
  np.testing.assert_allclose(np.concatenate(out), [1.5, 2.0], rtol=1e-6)         # d4pg/learning.py:235-237


# ------------------------------------------------------------------ optimizer and layers against PyTorch's own
def test_adam_form_matches_torch_adam():
  """eps_mode 0 (bias-corrected moments, eps outside the root) is torch.optim.Adam's update."""
  import torch
  rng = np.random.default_rng(10)
  p0 = rng.standard_normal(257).astype(f32)
  tp = torch.nn.Parameter(torch.tensor(p0))
  opt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
  adam = olearner.Adam(1e-3)
  params = {'w': p0.copy()}
  for _ in range(25):
    g = rng.standard_normal(257).astype(f32)
    tp.grad = torch.tensor(g)
    opt.step()
    params = adam.apply({'w': g}, params)
  np.testing.assert_allclose(params['w'], tp.detach().numpy(), rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize('size,k,s,out,before,after', [(84, 8, 4, 21, 2, 2), (21, 4, 2, 11, 1, 2), (11, 3, 1, 11, 1, 1),
                                                       (10, 3, 2, 5, 0, 1)])
def test_tf_same_padding_rule(size, k, s, out, before, after):
  # ceil(size / stride) outputs; the smaller half of the padding goes in front (TensorFlow's SAME)
  assert onets.tf_same_pad(size, k, s) == (out, before, after)
  from acme_b200 import networks
  assert networks.tf_same_pad(size, k, s) == (out, before)   # the product's copy of the rule


def test_oracle_conv_is_torch_conv_with_explicit_asymmetric_padding():
  import torch
  import torch.nn.functional as F
  rng = np.random.default_rng(11)
  x = torch.tensor(rng.standard_normal((2, 21, 21, 8)).astype(f32))
  w = torch.tensor(rng.standard_normal((4, 4, 8, 16)).astype(f32))          # HWIO like Sonnet
  b = torch.tensor(rng.standard_normal(16).astype(f32))
  y = onets._conv_same(x, w, b, 2)
  xp = F.pad(x.permute(0, 3, 1, 2), (1, 2, 1, 2))                            # total pad 3: 1 before, 2 after
  ref = F.conv2d(xp, w.permute(3, 2, 0, 1), b, stride=2).permute(0, 2, 3, 1)
  assert tuple(y.shape) == (2, 11, 11, 16)
  torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ the ring-of-steps table against materialised transitions
def _episodes(rng, n_eps, obs_dim=3):
  eps = []
  for _ in range(n_eps):
    T = int(rng.integers(2, 12))
    eps.append(dict(obs=rng.standard_normal((T + 1, obs_dim)).astype(f32), act=rng.integers(0, 4, T).astype(np.int32),
                    rew=rng.standard_normal(T).astype(f32), disc=np.concatenate([np.ones(T - 1, f32), np.zeros(1, f32)])))
  return eps


def test_ring_table_yields_the_reference_adders_transitions():
  """Feeding steps to the ring table and gathering every item gives exactly the (o, a, R, D, o') tuples that the reference
  adder's arithmetic (oracle.nstep, pinned to the reference's golden cases) materialises for the same episodes."""
  from oracle import nstep as onstep
  from oracle import replay as oreplay
  rng = np.random.default_rng(12)
  n, gamma = 3, 0.95
  table = oreplay.Table(4096, 8192, (3,), f32, (), np.int32, gamma, 0.6, max_window=n)
  expect = []
  for ep in _episodes(rng, 12):
    w = table.writer()
    T = len(ep['act'])
    window = []
    for t in range(T):
      table.append(w, ep['obs'][t], ep['act'][t], ep['rew'][t], ep['disc'][t], ep['obs'][t + 1])
      window.append(t)
      window = window[-n:]
      table.create_item(w, len(window), 1.0)
      expect.append((window[0], t + 1, list(window), ep))
    window = window[1:]
    while window:                                  # episode end: the shrinking tails (transition.py:147-172)
      table.create_item(w, len(window), 1.0)
      expect.append((window[0], T, list(window), ep))
      window = window[1:]
    table.close(w)
  assert table.size == len(expect)
  o, a, R, D, o2 = table.gather(np.arange(len(expect)))
  for i, (s, e, win, ep) in enumerate(expect):
    r, d = onstep.nstep_return([ep['rew'][t] for t in win], [ep['disc'][t] for t in win], f32(gamma))
    np.testing.assert_array_equal(o[i], ep['obs'][s])
    np.testing.assert_array_equal(o2[i], ep['obs'][e])
    assert a[i] == ep['act'][s] and R[i] == r and D[i] == d


def test_ring_table_fifo_and_stale_keys():
  from oracle import replay as oreplay
  table = oreplay.Table(8, 64, (1,), f32, (), np.int32, 0.99, 1.0, max_window=1)
  w = table.writer()
  keys = []
  for t in range(20):
    table.append(w, [t], 0, 1.0, 1.0, [t + 1])
    keys.append(table.create_item(w, 1, float(t + 1)))
  assert table.size == 8 and keys == list(range(20))                      # Fifo remover, max_size 8
  live = table.tree.leaves[:8]
  np.testing.assert_array_equal(np.sort(live), np.arange(13, 21, dtype=f32))   # priorities of items 12..19 (alpha = 1)
  before = table.tree.leaves.copy()
  table.update_priorities([3, 11], [100.0, 100.0])                        # evicted keys: ignored, like Reverb
  np.testing.assert_array_equal(table.tree.leaves, before)
  table.update_priorities([19, 19], [5.0, 7.0])                           # duplicates: the last one wins
  assert table.tree.leaves[19 % 8] == f32(7.0)
  k, pos, prob = table.sample(np.linspace(0.01, 0.99, 50).astype(f32), stratified=False)
  assert set(k.tolist()) <= set(range(12, 20))
  np.testing.assert_allclose(prob, table.tree.leaves[pos] / table.tree.total, rtol=1e-6)
