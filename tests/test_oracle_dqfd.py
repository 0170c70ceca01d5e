"""oracle.dqfd (CPU): `_n_step_transition_from_episode` (`acme/agents/tf/dqfd/agent.py:160-217`) against an independent
exact evaluation, and the draw rule's ranges.  Parity of this oracle is UNPINNED (no vectors in the reference, TF absent);
these are closed-form checks."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import dqfd


def _exact(rewards, discounts, n_step, g, first):
  """The docstring formula of agent.py:201-207 in exact rational arithmetic."""
  max_index = len(rewards) - 1
  last = min(first + n_step, max_index)
  r, c = Fraction(0), Fraction(1)
  for j in range(last - first):
    if j > 0:
      c *= Fraction(float(discounts[first + j - 1]))
    d = c * Fraction(g) ** j
    r += Fraction(float(rewards[first + 1 + j])) * d
  return r, d, last


@pytest.mark.parametrize('n_step', [1, 2, 5, 40])
def test_transition_matches_exact_evaluation(n_step):
  rng = np.random.default_rng(n_step)
  L = 23
  obs = np.arange(L * 2, dtype=np.float32).reshape(L, 2)
  act = np.arange(L, dtype=np.int32)
  rew = rng.integers(-4, 5, L).astype(np.float32)          # small integers, powers of two: every fp32 op is exact
  disc = rng.choice(np.array([0.5, 1.0], np.float32), L)
  disc[7] = 0.
  for first in range(0, L - 2):
    o, a, r, d, o2 = dqfd.n_step_transition_from_episode(obs, act, rew, disc, n_step, 0.5, first)
    want_r, want_d, last = _exact(rew, disc, n_step, 0.5, first)
    assert r.dtype == np.float32 and d.dtype == np.float32
    if n_step <= 5:                                            # beyond that 0.5^j * ints may round; short windows are exact
      assert Fraction(float(r)) == want_r and Fraction(float(d)) == want_d, (first, r, want_r, d, want_d)
    else:
      np.testing.assert_allclose(float(r), float(want_r), rtol=1e-6, atol=1e-6)
      np.testing.assert_allclose(float(d), float(want_d), rtol=1e-6)
    np.testing.assert_array_equal(o, obs[first])
    np.testing.assert_array_equal(o2, obs[last])
    assert a == act[first]
  with pytest.raises(AssertionError):
    dqfd.n_step_transition_from_episode(obs, act, rew, disc, n_step, 0.5, L - 2)   # maxval = max_index - 1 is exclusive


def test_last_transition_includes_the_final_reward_only_at_the_end():
  # agent.py:201-203: rewards are shifted by one so that last == max_index takes in the episode's last reward
  rew = np.array([9., 1., 2., 4.], np.float32)
  disc = np.ones(4, np.float32)
  obs, act = np.arange(4), np.arange(4)
  _, _, r, d, o2 = dqfd.n_step_transition_from_episode(obs, act, rew, disc, 3, 1.0, 0)
  assert r == 7. and d == 1. and o2 == 3            # rewards[1:4]; rewards[0] is ignored
  _, _, r, d, o2 = dqfd.n_step_transition_from_episode(obs, act, rew, disc, 3, 1.0, 1)
  assert r == 6. and o2 == 3                        # window cut at the episode's end: rewards[2:4]


def test_draw_ranges_and_frequencies():
  rng = np.random.default_rng(0)
  lengths = [3, 10, 4, 25]
  n, hits = 20000, 0
  firsts = {e: set() for e in range(4)}
  for u in rng.random((n, 3)).astype(np.float32):
    demo, e, first = dqfd.draw(u, 0.3, lengths)
    assert demo == (u[0] < np.float32(0.3))
    if demo:
      hits += 1
      assert 0 <= e < 4 and 0 <= first <= lengths[e] - 3
      firsts[e].add(first)
  assert abs(hits / n - 0.3) < 0.02
  assert all(firsts[e] == set(range(lengths[e] - 2)) for e in range(4))   # U{0 .. max_index - 2}, every value reached
  assert dqfd.draw((0.3, 0.5, 0.5), 0.3, lengths)[0] is False             # ratio itself: replay
  assert dqfd.draw((0.0, 0.99999994, 0.99999994), 0.3, lengths) == (True, 3, 22)


def test_mix_replaces_only_the_chosen_rows():
  rng = np.random.default_rng(1)
  B = 12
  eps = [(rng.standard_normal((6, 3)).astype(np.float32), rng.integers(0, 4, 6).astype(np.int32),
          rng.standard_normal(6).astype(np.float32), np.ones(6, np.float32)) for _ in range(3)]
  batch = dict(o_tm1=np.zeros((B, 3), np.float32), a_tm1=np.full(B, -1, np.int32), R=np.zeros(B, np.float32),
               D=np.zeros(B, np.float32), o_t=np.zeros((B, 3), np.float32), keys=np.arange(B, dtype=np.uint64),
               prob=np.full(B, 0.25, np.float32))
  u3 = rng.random((B, 3)).astype(np.float32)
  out, is_demo = dqfd.mix(batch, eps, u3, 0.5, 3, 0.99)
  assert 0 < is_demo.sum() < B
  assert (out['keys'][is_demo] == dqfd.NO_KEY).all() and (out['prob'][is_demo] == 1.).all()
  assert (out['keys'][~is_demo] == np.arange(B, dtype=np.uint64)[~is_demo]).all() and (out['prob'][~is_demo] == 0.25).all()
  assert (out['a_tm1'][~is_demo] == -1).all() and (out['a_tm1'][is_demo] >= 0).all()


def test_pack_episodes_host_logic():
  """The product's host half (no GPU): rows packed like the table's batch rows, offsets, and its argument checks."""
  import helpers
  from acme_b200 import dqfd as product
  spec, table, server, adder, oracle = helpers.make_pair((5,), np.float32, 4, 3, 0.99, 0.6, max_size=50)
  rng = np.random.default_rng(0)
  eps = [(rng.standard_normal((L, 5)).astype(np.float32), rng.integers(0, 4, L).astype(np.int32),
          rng.standard_normal(L).astype(np.float32), np.ones(L, np.float32)) for L in (3, 7, 4)]
  obs, act, rew, disc, off = product.pack_episodes(eps, table)
  assert obs.shape == (14, 20) and obs.dtype == np.uint8 and act.shape == (14, 4)
  np.testing.assert_array_equal(off, [0, 3, 10, 14])
  np.testing.assert_array_equal(obs.view(np.float32)[3:10], eps[1][0])
  np.testing.assert_array_equal(act.view(np.int32)[:, 0][10:], eps[2][1])
  np.testing.assert_array_equal(rew[:3], eps[0][2])
  with pytest.raises(ValueError, match='at least 3 steps'):
    product.pack_episodes([tuple(x[:2] for x in eps[0])], table)
  with pytest.raises(ValueError, match='no demonstration episodes'):
    product.pack_episodes([], table)
  with pytest.raises(ValueError, match='same length'):
    product.pack_episodes([(eps[0][0], eps[0][1], eps[0][2], np.ones(5, np.float32))], table)
  server.stop()
