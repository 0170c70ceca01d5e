"""Adders: the actor-side insert seam (`acme/adders/base.py:24-82`).

`SequenceAdder` (`acme/adders/reverb/sequence.py:29-127`) uses the same ring: every step is appended once
and an item is "the last `sequence_length` steps"; the episode's final step and its zero padding are appended as
ordinary steps, and K3's sequence form (`b200rl_replay_gather_sequences`) copies the T rows of a sampled item.

`NStepTransitionAdder` keeps the reference's constructor, `add_first / add / reset` protocol,
error behaviour and item bookkeeping (`acme/adders/reverb/base.py:62-176`,
`acme/adders/reverb/transition.py:87-190`), but it does NOT materialise transitions on the host:
each environment step is appended once to the HBM step ring, and an item is just "the last
`len(buffer)` steps" (Reverb's `create_item(num_timesteps=...)`).  The n-step return and discount
of an item are built on the GPU at gather time (K3) with the reference's fp32 arithmetic.
"""

from __future__ import annotations

import abc
import collections
from typing import Callable, Mapping, NamedTuple, Optional

import numpy as np

from acme_b200 import specs, tree
from acme_b200.replay import DEFAULT_PRIORITY_TABLE


class Adder(abc.ABC):
  """`acme/adders/base.py:24-82`."""

  @abc.abstractmethod
  def add_first(self, timestep):
    ...

  @abc.abstractmethod
  def add(self, action, next_timestep, extras=()):
    ...

  @abc.abstractmethod
  def reset(self):
    ...


class Step(NamedTuple):
  """`acme/adders/reverb/base.py:33-40`."""
  observation: object
  action: object
  reward: object
  discount: object
  start_of_episode: object
  extras: object


class PriorityFnInput(NamedTuple):
  """`acme/adders/reverb/base.py:43-50`."""
  observations: object
  actions: object
  rewards: object
  discounts: object
  start_of_episode: object
  extras: object


PriorityFn = Callable[[PriorityFnInput], float]
PriorityFnMapping = Mapping[str, PriorityFn]


def _default_priority(_) -> float:
  return 1.


class ReverbAdder(Adder):
  """Episode state machine shared by the adders (`acme/adders/reverb/base.py:62-176`)."""

  def __init__(self, client, buffer_size: int, max_sequence_length: int, delta_encoded: bool = False,
               chunk_length: Optional[int] = None, priority_fns: Optional[PriorityFnMapping] = None):
    if priority_fns:
      self._priority_fns = dict(priority_fns)
      self._default_priorities = False
    else:
      self._priority_fns = {DEFAULT_PRIORITY_TABLE: _default_priority}
      self._default_priorities = True
    self._client = client
    self._max_sequence_length = max_sequence_length
    self._delta_encoded = delta_encoded
    self._chunk_length = chunk_length
    self.__writer = None
    self._buffer = collections.deque(maxlen=buffer_size)
    self._next_observation = None
    self._start_of_episode = False

  @property
  def _writer(self):
    if self.__writer is None:  # created lazily, one per episode
      self.__writer = self._client.writer(self._max_sequence_length, delta_encoded=self._delta_encoded,
                                          chunk_length=self._chunk_length)
    return self.__writer

  def add_priority_table(self, table_name: str, priority_fn: PriorityFn):
    if table_name in self._priority_fns:
      raise ValueError('A priority function already exists for {}.'.format(table_name))
    self._priority_fns[table_name] = priority_fn
    self._default_priorities = False

  def reset(self):
    if self.__writer:
      self.__writer.close()
      self.__writer = None
    self._buffer.clear()
    self._next_observation = None

  def add_first(self, timestep):
    if not timestep.first():
      raise ValueError('adder.add_first with an initial timestep (i.e. one for which '
                       'timestep.first() is True')
    if self._next_observation is not None:
      raise ValueError('adder.reset must be called before adder.add_first (called automatically '
                       'if `next_timestep.last()` is true when `add` is called).')
    self._next_observation = timestep.observation
    self._start_of_episode = True

  def add(self, action, next_timestep, extras=()):
    if self._next_observation is None:
      raise ValueError('adder.add_first must be called before adder.add.')
    step = Step(observation=self._next_observation, action=action, reward=next_timestep.reward,
                discount=next_timestep.discount, start_of_episode=self._start_of_episode, extras=extras)
    self._buffer.append(step)
    self._append_step(step, next_timestep.observation)
    self._next_observation = next_timestep.observation
    self._start_of_episode = False
    self._write()
    if next_timestep.last():
      self._write_last()
      self.reset()

  @abc.abstractmethod
  def _append_step(self, step: Step, next_observation):
    ...

  @abc.abstractmethod
  def _write(self):
    ...

  @abc.abstractmethod
  def _write_last(self):
    ...


def _stack(values):
  return tree.map_structure(lambda *xs: np.asarray(xs), *values)


class NStepTransitionAdder(ReverbAdder):
  """`acme/adders/reverb/transition.py:38-190`."""

  def __init__(self, client, n_step: int, discount: float,
               priority_fns: Optional[PriorityFnMapping] = None):
    if n_step < 1:
      raise ValueError('n_step must be at least 1')
    # float32 like the reference (transition.py:111); the replay table applies it on the GPU and
    # must have been created with the same value.
    self._discount = np.float32(discount)
    self.n_step = n_step
    super().__init__(client=client, buffer_size=n_step, max_sequence_length=1, priority_fns=priority_fns)
    for name in self._priority_fns:
      self._check_table(name)

  def _check_table(self, name):
    server = getattr(self._client, 'server', None)
    if server is None or name not in getattr(server, 'tables', {}):
      return
    t = server.tables[name]
    if np.float32(t.discount) != self._discount:
      raise ValueError(f'table {name!r} was built with discount {t.discount}, adder uses {self._discount}')
    if t.max_window < self.n_step:
      raise ValueError(f'table {name!r} supports windows up to {t.max_window} < n_step {self.n_step}')

  def _append_step(self, step: Step, next_observation):
    self._writer.append_step(step.observation, step.action, step.reward, step.discount,
                             next_observation, extras=step.extras, tables=list(self._priority_fns))

  def _priorities(self):
    if self._default_priorities:
      return {t: 1. for t in self._priority_fns}
    # user priority functions see the stacked window + a zero terminal step (utils.py:52-104)
    first = self._buffer[0]
    zeros = lambda x: tree.map_structure(lambda v: np.zeros_like(np.asarray(v)), x)
    final = Step(self._next_observation, zeros(first.action), zeros(first.reward), zeros(first.discount),
                 False, zeros(first.extras))
    steps = list(self._buffer) + [final]
    fn_input = PriorityFnInput(*[_stack([s[i] for s in steps]) for i in range(6)])
    return {t: float(fn(fn_input)) for t, fn in self._priority_fns.items()}

  def _write(self):
    # the item spans the whole deque (transition.py:120-145): the last len(buffer) appended steps
    for table, priority in self._priorities().items():
      self._writer.create_item(table=table, num_timesteps=len(self._buffer), priority=priority)

  def _write_last(self):
    self._buffer.popleft()
    while self._buffer:
      self._write()
      self._buffer.popleft()

  @classmethod
  def signature(cls, environment_spec: specs.EnvironmentSpec, extras_spec=()):
    """(obs, action, reward, discount, next_obs[, extras]) specs (transition.py:174-190)."""
    sig = [environment_spec.observations, environment_spec.actions, environment_spec.rewards,
           environment_spec.discounts, environment_spec.observations]
    if extras_spec:
      sig.append(extras_spec)
    return tuple(sig)


class SequenceAdder(ReverbAdder):
  """`acme/adders/reverb/sequence.py:29-127`: fixed-length (possibly overlapping) sequences for recurrent learners."""

  def __init__(self, client, sequence_length: int, period: int, delta_encoded: bool = False,
               chunk_length: Optional[int] = None, priority_fns: Optional[PriorityFnMapping] = None,
               pad_end_of_episode: bool = True):
    super().__init__(client=client, buffer_size=sequence_length, max_sequence_length=sequence_length,
                     delta_encoded=delta_encoded, chunk_length=chunk_length, priority_fns=priority_fns)
    self._period = period
    self._step = 0
    self._pad_end_of_episode = pad_end_of_episode
    for name in self._priority_fns:
      server = getattr(self._client, 'server', None)
      t = getattr(server, 'tables', {}).get(name) if server is not None else None
      if t is not None and t.max_window < sequence_length:
        raise ValueError(f'table {name!r} supports windows up to {t.max_window} < sequence_length {sequence_length}')

  def reset(self):
    self._step = 0
    super().reset()

  def _append_step(self, step: Step, next_observation):
    self._writer.append_step(step.observation, step.action, step.reward, step.discount, next_observation,
                             extras=step.extras, tables=list(self._priority_fns),
                             start_of_episode=step.start_of_episode)

  def _write(self):
    # the step itself went into the ring in _append_step (sequence.py:77-81)
    self._step += 1
    self._maybe_add_priorities()

  def _write_last(self):
    # sequence.py:83-108: the final observation becomes a step of its own with zero action / reward / discount /
    # extras, then (optionally) all-zero steps up to the next point at which a sequence is due
    first = self._buffer[0]
    zeros = lambda x: tree.map_structure(lambda v: np.zeros_like(np.asarray(v)), x)
    final = Step(self._next_observation, zeros(first.action), zeros(first.reward), zeros(first.discount), False,
                 zeros(first.extras))
    zero_obs = zeros(self._next_observation)
    self._buffer.append(final)
    self._append_step(final, zero_obs)
    self._step += 1
    if self._pad_end_of_episode:
      zero_step = final._replace(observation=zero_obs)
      if self._step <= self._max_sequence_length:
        padding = self._max_sequence_length - self._step
      else:
        padding = self._period - (self._step - self._max_sequence_length)
      for _ in range(padding):
        self._buffer.append(zero_step)
        self._append_step(zero_step, zero_obs)
        self._step += 1
    self._maybe_add_priorities()

  def _maybe_add_priorities(self):
    # sequence.py:110-127
    L = self._max_sequence_length
    if not (self._step == L or (self._step > L and (self._step - L) % self._period == 0)):
      return
    steps = list(self._buffer)
    if self._default_priorities:
      priorities = {t: 1. for t in self._priority_fns}
    else:
      fn_input = PriorityFnInput(*[_stack([s[i] for s in steps]) for i in range(6)])
      priorities = {t: float(fn(fn_input)) for t, fn in self._priority_fns.items()}
    for table, priority in priorities.items():
      self._writer.create_item(table=table, num_timesteps=len(steps), priority=priority)

  @classmethod
  def signature(cls, environment_spec: specs.EnvironmentSpec, extras_spec=()):
    """The per-step spec of a sequence table: a `Step` of specs (what later Acme versions' SequenceAdder.signature
    returns, minus the leading time axis)."""
    return Step(observation=environment_spec.observations, action=environment_spec.actions,
                reward=environment_spec.rewards, discount=environment_spec.discounts,
                start_of_episode=specs.Array((), np.bool_), extras=extras_spec)
