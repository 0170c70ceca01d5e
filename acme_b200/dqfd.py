"""DQfD: a DQN learner whose batches mix replay with demonstrations (`acme/agents/tf/dqfd/agent.py:37-217`).

The reference builds `sample_from_datasets([replay dataset, demonstrations.map(_n_step_transition_from_episode)],
[1 - ratio, ratio])` and hands the result to the unchanged `DQNLearner`.  Here the demonstration episodes are uploaded
once into flat HBM arrays (`DemonstrationSet`) and `MixedReplayDataset` -- a `ReplayDataset` -- overwrites, after K1 /
K3 have filled the batch from replay, the rows whose draw falls below `ratio` with demonstration transitions
(`csrc/dqfd.cu`): no host decision, so the learner's captured step stays one CUDA graph.  The learner is `dqn.DQNLearner`
itself, as in the reference.
"""

from __future__ import annotations

from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

from acme_b200 import _capi, actors, adders, agent, dqn, loggers, replay, specs, tree


class DemonstrationSet:
  """Whole demonstration episodes in HBM.  `episodes`: iterable of (observations [L, ...], actions [L, ...], rewards [L],
  discounts [L]) -- what the reference's `demonstration_dataset` yields (`dqfd/agent.py:160-183`: the first reward and
  discount and the last action of an episode are ignored).  Rows are packed exactly like the table's batch rows."""

  def __init__(self, episodes: Iterable[Tuple], table: replay.Table):
    import torch
    obs, act, rew, disc, offsets = pack_episodes(episodes, table)
    dev = torch.device('cuda', table.device)
    self.obs_bytes, self.act_bytes = obs.shape[1], act.shape[1]
    self.obs, self.act = torch.from_numpy(obs).to(dev), torch.from_numpy(act).to(dev)
    self.rew, self.disc = torch.from_numpy(rew).to(dev), torch.from_numpy(disc).to(dev)
    self.offsets = torch.from_numpy(offsets).to(dev)
    self.num_episodes = len(offsets) - 1
    self.num_steps = int(offsets[-1])


def pack_episodes(episodes: Iterable[Tuple], table: replay.Table):
  """Host half of `DemonstrationSet`: (obs rows u8 [steps, obs_bytes], act rows u8 [steps, act_bytes], rewards f32 [steps],
  discounts f32 [steps], episode offsets i64 [episodes + 1]) with the table's own row packing."""
  obs_rows, act_rows, rew, disc, offsets = [], [], [], [], [0]
  act_leaves = (lambda a: (a, ())) if table.has_extras else (lambda a: a)
  for observations, actions, rewards, discounts in episodes:
    L = len(rewards)
    if L < 3:
      raise ValueError(f'a demonstration episode needs at least 3 steps, got {L}')   # first ~ U{0 .. L - 3}
    if len(discounts) != L or len(tree.flatten(actions)[0]) != L or len(tree.flatten(observations)[0]) != L:
      raise ValueError('observations, actions, rewards and discounts of an episode must have the same length')
    for i in range(L):
      obs_rows.append(np.array(table.obs_packer.pack(tree.map_structure(lambda x: x[i], observations))))
      act_rows.append(np.array(table.act_packer.pack(act_leaves(tree.map_structure(lambda x: x[i], actions)))))
    rew.append(np.asarray(rewards, np.float32))
    disc.append(np.asarray(discounts, np.float32))
    offsets.append(offsets[-1] + L)
  if len(offsets) < 2:
    raise ValueError('no demonstration episodes')
  pad = lambda rows, n: np.stack([np.pad(r, (0, n - r.size)) for r in rows])   # single-leaf rows carry no row padding
  return (pad(obs_rows, max(table.obs_packer.nbytes, 1)), pad(act_rows, max(table.act_packer.nbytes, 1)),
          np.concatenate(rew), np.concatenate(disc), np.asarray(offsets, np.int64))


class MixedReplayDataset(replay.ReplayDataset):
  """`sample_from_datasets([replay, demonstrations], [1 - ratio, ratio])` over batch elements (`dqfd/agent.py:111-122`).
  Three uniforms per element: source, episode, first step -- a device Philox stream keyed by (seed, draw counter) like
  K1's own draws, or injected with `inject_demo_uniforms` (tests).  Demonstration rows carry the reference's constant
  SampleInfo (probability 1; their key names no item, so the learner's priority write-back skips them)."""

  # K3 cannot write conv1's row image for rows that are replaced afterwards: the learner converts the mixed batch instead
  supports_gather_rows = False
  _SEED_SALT = 0x5DEECE66D

  def __init__(self, table: replay.Table, batch_size: int, demonstrations, demonstration_ratio: float, n_step: int,
               discount: float, seed: int = 0, stratified: bool = True):
    import torch
    super().__init__(table, batch_size, seed=seed, stratified=stratified)
    if table.is_sequence:
      raise ValueError('demonstration mixing is defined for transition tables')
    if not 0. <= demonstration_ratio <= 1.:
      raise ValueError('demonstration_ratio must be in [0, 1]')
    self.demos = demonstrations if isinstance(demonstrations, DemonstrationSet) else DemonstrationSet(demonstrations, table)
    self.ratio, self.n_step, self.discount = float(demonstration_ratio), int(n_step), float(np.float32(discount))
    dev = self.R.device
    self.u3 = torch.zeros(3 * self.B, dtype=torch.float32, device=dev)
    self.is_demo = torch.zeros(self.B, dtype=torch.int32, device=dev)
    self._held = False

  def inject_demo_uniforms(self, uniforms3):
    """The next batch uses these 3B draws ([B, 3]: source, episode, first step) instead of the Philox stream."""
    self.u3.copy_(uniforms3.reshape(-1))
    self._held = True

  def sample_only(self, uniforms=None, bump: bool = True):
    if self._held:
      self._held = False
    else:   # before the parent call: it may advance the draw counter
      _capi.call('b200rl_uniform', _capi.ptr(self.u3), 3 * self.B, self.seed ^ self._SEED_SALT, _capi.ptr(self.counter), 0,
                 _capi.current_stream())
    super().sample_only(uniforms, bump)

  def gather_only(self, rows=None):
    if rows is not None:
      raise ValueError('MixedReplayDataset does not write row images (supports_gather_rows is False)')
    super().gather_only(None)
    d = self.demos
    _capi.call('b200rl_demo_mix', self.B, _capi.ptr(d.obs), d.obs_bytes, _capi.ptr(d.act), d.act_bytes, _capi.ptr(d.rew),
               _capi.ptr(d.disc), _capi.ptr(d.offsets), d.num_episodes, self.n_step, self.discount, _capi.ptr(self.u3),
               self.ratio, _capi.ptr(self.o_tm1), _capi.ptr(self.a_tm1), _capi.ptr(self.R), _capi.ptr(self.D),
               _capi.ptr(self.o_t), _capi.ptr(self.keys), _capi.ptr(self.prob), _capi.ptr(self.is_demo), _capi.current_stream())


class DQfD(agent.Agent):
  """`acme/agents/tf/dqfd/agent.py:37-157`: DQN whose learner batches contain demonstrations with probability
  `demonstration_ratio`.  The replay table is Uniform, as in the reference (`:92-98`)."""

  def __init__(self, environment_spec: specs.EnvironmentSpec, network, demonstration_dataset, demonstration_ratio: float,
               batch_size: int = 256, prefetch_size: int = 4, target_update_period: int = 100,
               samples_per_insert: float = 32.0, min_replay_size: int = 1000, max_replay_size: int = 1000000,
               importance_sampling_exponent: float = 0.2, n_step: int = 5, epsilon: Optional[float] = None,
               learning_rate: float = 1e-3, discount: float = 0.99, logger: loggers.Logger = None, seed: int = 0,
               use_cuda_graph: bool = True, slot_capacity: Optional[int] = None):
    table = replay.Table(
        name=replay.DEFAULT_PRIORITY_TABLE, sampler=replay.selectors.Uniform(), remover=replay.selectors.Fifo(),
        max_size=max_replay_size, rate_limiter=replay.rate_limiters.MinSize(1),
        signature=adders.NStepTransitionAdder.signature(environment_spec), max_window=max(n_step, 1), discount=discount,
        device=network.device, slot_capacity=slot_capacity)
    self._server = replay.Server([table], port=None)
    address = f'localhost:{self._server.port}'
    adder = adders.NStepTransitionAdder(client=replay.Client(address), n_step=n_step, discount=discount)
    replay_client = replay.TFClient(address)
    dataset = MixedReplayDataset(table, batch_size, demonstration_dataset, demonstration_ratio, n_step=n_step,
                                 discount=discount, seed=seed)
    policy = actors.EpsilonGreedyPolicy(network, 0.05 if epsilon is None else epsilon, seed=seed)
    target_network = network.clone()
    actor = actors.FeedForwardActor(policy, adder)
    learner = dqn.DQNLearner(network=network, target_network=target_network, discount=discount,
                             importance_sampling_exponent=importance_sampling_exponent, learning_rate=learning_rate,
                             target_update_period=target_update_period, dataset=dataset, replay_client=replay_client,
                             logger=logger, checkpoint=False, use_cuda_graph=use_cuda_graph)
    self._learner_obj = learner
    self._table = table
    self._dataset = dataset
    super().__init__(actor=actor, learner=learner, min_observations=max(batch_size, min_replay_size),
                     observations_per_step=float(batch_size) / samples_per_insert)
