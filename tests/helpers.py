"""Shared drivers: feed the same synthetic episodes to the CUDA replay (through the product adder)
and to the oracle table, so both hold the same items under the same keys."""
import numpy as np

from acme_b200 import adders, dm_env, replay, specs
from oracle import nstep as onstep
from oracle import replay as oreplay


def make_pair(obs_shape=(8, 8, 4), obs_dtype=np.uint8, num_actions=4, n_step=3, discount=0.99, alpha=0.6,
              max_size=500, slot_capacity=None, stage_slots=0, act_dim=None, frame_stack=0):
  """num_actions discrete actions, or (act_dim given) a float32 action vector in [-1, 1]."""
  aspec = (specs.DiscreteArray(num_actions) if act_dim is None else
           specs.BoundedArray((act_dim,), np.float32, -1., 1.))
  spec = specs.EnvironmentSpec(specs.Array(obs_shape, obs_dtype), aspec,
                               specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
  table = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Prioritized(alpha), replay.selectors.Fifo(),
                       max_size=max_size, rate_limiter=replay.rate_limiters.MinSize(1),
                       signature=adders.NStepTransitionAdder.signature(spec), max_window=n_step,
                       discount=discount, slot_capacity=slot_capacity, stage_slots=stage_slots, frame_stack=frame_stack)
  server = replay.Server([table])
  adder = adders.NStepTransitionAdder(replay.Client(server), n_step=n_step, discount=discount)
  oracle = oreplay.Table(max_size, table.slot_capacity, obs_shape, obs_dtype, () if act_dim is None else (act_dim,),
                         np.int32 if act_dim is None else np.float32, discount, alpha, max_window=n_step)
  return spec, table, server, adder, oracle


def random_obs(rng, shape, dtype):
  if np.dtype(dtype) == np.uint8:
    return rng.integers(0, 256, shape, dtype=np.uint8)
  return rng.standard_normal(shape).astype(dtype)


def feed_episode(rng, adder, oracle, T, n_step, obs_shape, obs_dtype, num_actions, terminal=True, act_dim=None):
  """One episode of T steps into both stores (oracle items follow SURVEY App. A.1)."""
  o = random_obs(rng, obs_shape, obs_dtype)
  adder.add_first(dm_env.restart(o))
  w = oracle.writer()
  for k in range(1, T + 1):
    a = np.int32(rng.integers(num_actions)) if act_dim is None else rng.uniform(-1, 1, act_dim).astype(np.float32)
    r = np.float32(rng.choice([-1., 0., 1., 0.5, 2.5]))
    last = k == T
    d = np.float32(0. if (last and terminal) else rng.choice([1., 1., 0.9]))
    o2 = random_obs(rng, obs_shape, obs_dtype)
    ts = dm_env.TimeStep(dm_env.StepType.LAST if last else dm_env.StepType.MID, r, d, o2)
    adder.add(a, ts)
    oracle.append(w, o, a, r, d, o2)
    oracle.create_item(w, min(k, n_step), 1.0)
    if last:
      m = min(n_step, T)
      for j in range(1, m):
        oracle.create_item(w, m - j, 1.0)
      oracle.close(w)
    o = o2


def feed_stacked_episode(rng, adder, oracle, T, n_step, frame_shape, num_frames, num_actions, terminal=True):
  """One episode whose observations are FrameStacker stacks (oracle.replay.FrameStacker <- frame_stacking.py:64-88) of
  random uint8 frames: what AtariWrapper hands to the adder.  The oracle table stores the stacks whole."""
  stacker = oreplay.FrameStacker(num_frames)
  o = stacker.step(rng.integers(0, 256, frame_shape, dtype=np.uint8))
  adder.add_first(dm_env.restart(o))
  w = oracle.writer()
  for k in range(1, T + 1):
    a = np.int32(rng.integers(num_actions))
    r = np.float32(rng.choice([-1., 0., 1., 0.5]))
    last = k == T
    d = np.float32(0. if (last and terminal) else 1.)
    o2 = stacker.step(rng.integers(0, 256, frame_shape, dtype=np.uint8))
    adder.add(a, dm_env.TimeStep(dm_env.StepType.LAST if last else dm_env.StepType.MID, r, d, o2))
    oracle.append(w, o, a, r, d, o2)
    oracle.create_item(w, min(k, n_step), 1.0)
    if last:
      m = min(n_step, T)
      for j in range(1, m):
        oracle.create_item(w, m - j, 1.0)
      oracle.close(w)
    o = o2


def sync_oracle_leaves(table, oracle):
  """Copies the GPU's leaf weights into the oracle tree (after checking they agree to 1 ulp) so that
  later bit-exact comparisons do not depend on libm-vs-CUDA pow()."""
  L = oracle.tree.L
  got = table.read_tree_level(L)[:oracle.tree.levels[L].shape[0]]
  want = oracle.tree.levels[L]
  np.testing.assert_allclose(got, want, rtol=2.5e-7, atol=0)
  oracle.tree.levels[L][:] = got
  oracle.tree.rebuild()


def sync_oracle_leaves_loose(table, oracle, rtol=5e-3, atol=1e-4):
  """After a learner step the new priorities are |td| computed by two fp32 pipelines that agree to
  ~1e-5, so leaves agree loosely; adopt the GPU's leaves so the next draw is compared bit-exactly."""
  L = oracle.tree.L
  got = table.read_tree_level(L)[:oracle.tree.levels[L].shape[0]]
  np.testing.assert_allclose(got, oracle.tree.levels[L], rtol=rtol, atol=atol)
  oracle.tree.levels[L][:] = got
  oracle.tree.rebuild()


def make_dqn_dp_learner(rank, world, group, peer_exchange, items=4096, precision=0):
  """Atari-shaped DQN learner on replay shard `rank` of a data-parallel job (same parameters on every rank)."""
  import sys, os
  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  import bench
  from acme_b200 import adders, dqn, loggers, networks, replay, specs
  spec = specs.EnvironmentSpec(specs.Array(bench.OBS_SHAPE, np.uint8), specs.DiscreteArray(bench.NUM_ACTIONS),
                               specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
  table = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Prioritized(0.6), replay.selectors.Fifo(),
                       max_size=items, rate_limiter=replay.rate_limiters.MinSize(1),
                       signature=adders.NStepTransitionAdder.signature(spec), max_window=3, discount=0.99, device=rank,
                       slot_capacity=items + 4096, shard_count=world, shard_rank=rank, stage_slots=4096)
  server = replay.Server([table])
  bench.fill_replay(table, items, 3, seed=77 + rank)
  net = networks.DQNAtariNetwork(bench.NUM_ACTIONS, device=rank, precision=precision, seed=5)
  ds = replay.ReplayDataset(table, 64, seed=11 + rank)
  learner = dqn.DQNLearner(net, net.clone(), 0.99, 0.2, 1e-3, 100, ds, replay_client=replay.Client(server),
                           logger=loggers.NoOpLogger(), process_group=group, peer_exchange=peer_exchange)
  learner._keepalive = (server, table)
  return learner
