"""DQfD demonstration mixing, CPU restatement.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows `acme/agents/tf/dqfd/agent.py`:
  * `:160-217` `_n_step_transition_from_episode`: one n-step transition out of a whole demonstration episode, with ITS
    arithmetic (a cumulative product of the environment discounts times powers of the agent discount, then one sum) --
    not the adder's running update (`acme/adders/reverb/transition.py:135-145`), so the two may differ in the last ulp;
  * `:111-122` `sample_from_datasets([replay, demonstrations], [1 - ratio, ratio])`: every batch ELEMENT comes from the
    demonstrations with probability `ratio`.

Parity UNPINNED: TensorFlow is not installable here and the reference holds no test or vector for this file.  The random
draws (`tf.random.uniform` for `first`, the dataset sampler's own generator for the source choice) are replaced by
explicit uniforms in [0, 1): source = demonstrations iff u0 < ratio, episode = floor(u1 * num_episodes),
first = floor(u2 * (max_index - 1)) -- the integer `tf.random.uniform(minval=0, maxval=max_index - 1)` draws, U{0 ..
max_index - 2}.
"""
import numpy as np

NO_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)   # the reference's key 0 names no Reverb item; here keys count from 0, so "no item" is ~0


def n_step_transition_from_episode(observations, actions, rewards, discounts, n_step, discount, first):
  """agent.py:184-219 for a given `first`.  Arrays have the episode's length L on axis 0; the first reward / discount and
  the last action are ignored, exactly as the reference says.  Returns (o_t, a_t, r_t, d_t, o_tp1) with fp32 r_t, d_t."""
  rewards = np.asarray(rewards, np.float32)
  discounts = np.asarray(discounts, np.float32)
  max_index = rewards.shape[0] - 1                                           # :184
  assert 0 <= first <= max_index - 2, (first, max_index)                       # :185-186 (maxval is exclusive)
  last = min(first + n_step, max_index)                                      # :187
  m = last - first
  g = float(np.float32(discount))                                            # the Python float becomes an fp32 tensor
  additional = np.array([np.float32(g ** j) for j in range(m)], np.float32)   # :194-196: g^0 .. g^(m-1), fp32 (correctly rounded)
  cum = np.concatenate([np.ones(1, np.float32), np.cumprod(discounts[first:last - 1], dtype=np.float32)])   # :198
  disc = (cum * additional).astype(np.float32)                               # :200
  terms = (rewards[first + 1:last + 1] * disc).astype(np.float32)            # :204
  r = np.float32(0.)
  for t in terms:                                                            # reduce_sum in index order
    r = np.float32(r + t)
  return observations[first], actions[first], r, disc[-1], observations[last]


def draw(u3, ratio, episode_lengths):
  """The three uniforms of one batch element -> (is_demo, episode, first)."""
  u0, u1, u2 = (np.float32(x) for x in u3)
  if not u0 < np.float32(ratio):
    return False, -1, -1
  E = len(episode_lengths)
  e = min(max(int(np.float32(u1 * np.float32(E))), 0), E - 1)
  max_index = int(episode_lengths[e]) - 1
  first = min(max(int(np.float32(u2 * np.float32(max_index - 1))), 0), max_index - 2)
  return True, e, first


def mix(batch, episodes, uniforms3, ratio, n_step, discount):
  """batch = dict(o_tm1, a_tm1, R, D, o_t, keys, prob) of NumPy arrays sampled from replay (modified in place and
  returned); episodes = list of (observations, actions, rewards, discounts).  Rows chosen for the demonstrations are
  replaced; their SampleInfo is the reference's constant one (probability 1, agent.py:208-217)."""
  B = batch['R'].shape[0]
  u = np.asarray(uniforms3, np.float32).reshape(B, 3)
  lengths = [len(ep[2]) for ep in episodes]
  is_demo = np.zeros(B, bool)
  for b in range(B):
    demo, e, first = draw(u[b], ratio, lengths)
    if not demo:
      continue
    is_demo[b] = True
    o, a, r, d, o2 = n_step_transition_from_episode(*episodes[e], n_step=n_step, discount=discount, first=first)
    batch['o_tm1'][b], batch['a_tm1'][b], batch['R'][b], batch['D'][b], batch['o_t'][b] = o, a, r, d, o2
    batch['keys'][b] = NO_KEY
    batch['prob'][b] = 1.0
  return batch, is_demo
