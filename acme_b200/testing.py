"""Test doubles the reference's wiring tests rely on (`acme/testing/fakes.py:36-219`): a counting
Actor and fake environments that emit `spec.generate_value()`."""

from typing import Sequence

import numpy as np

from acme_b200 import core, dm_env, specs, tree


def _generate(spec):
  return tree.map_structure(lambda s: s.generate_value(), spec)


class Actor(core.Actor):
  """Validates what it is given against the spec and counts updates (`fakes.py:36-61`)."""

  def __init__(self, spec: specs.EnvironmentSpec):
    self._spec = spec
    self.num_updates = 0

  def select_action(self, observation):
    tree.map_structure(lambda s, v: s.validate(v), self._spec.observations, observation)
    return _generate(self._spec.actions)

  def observe_first(self, timestep):
    tree.map_structure(lambda s, v: s.validate(v), self._spec.observations, timestep.observation)

  def observe(self, action, next_timestep):
    tree.map_structure(lambda s, v: s.validate(v), self._spec.actions, action)
    tree.map_structure(lambda s, v: s.validate(v), self._spec.observations, next_timestep.observation)

  def update(self):
    self.num_updates += 1


class Environment(dm_env.Environment):
  """`fakes.py:80-144`."""

  def __init__(self, spec: specs.EnvironmentSpec, *, episode_length: int = 25):
    d = spec.discounts
    if not isinstance(d, specs.BoundedArray) or not np.isclose(d.minimum, 0) or not np.isclose(d.maximum, 1):
      raise ValueError('discount_spec must be a BoundedArray in [0, 1].')
    self._spec, self._episode_length, self._step = spec, episode_length, 0

  def reset(self):
    self._step = 1
    return dm_env.restart(_generate(self._spec.observations))

  def step(self, action):
    if not self._step:
      return self.reset()
    tree.map_structure(lambda s, v: s.validate(v), self._spec.actions, action)
    obs, rew, disc = _generate(self._spec.observations), _generate(self._spec.rewards), _generate(self._spec.discounts)
    if self._episode_length and self._step == self._episode_length:
      self._step = 0
      return dm_env.TimeStep(dm_env.StepType.LAST, rew, disc, obs)
    self._step += 1
    return dm_env.transition(reward=rew, observation=obs, discount=disc)

  def action_spec(self):
    return self._spec.actions

  def observation_spec(self):
    return self._spec.observations

  def reward_spec(self):
    return self._spec.rewards

  def discount_spec(self):
    return self._spec.discounts


class DiscreteEnvironment(Environment):
  """`fakes.py:147-175`."""

  def __init__(self, *, num_actions: int = 1, num_observations: int = 1, action_dtype=np.int32, obs_dtype=np.int32,
               reward_dtype=np.float32, obs_shape: Sequence[int] = (), **kwargs):
    super().__init__(spec=specs.EnvironmentSpec(
        observations=specs.BoundedArray(obs_shape, obs_dtype, obs_dtype(0), obs_dtype(num_observations - 1)),
        actions=specs.DiscreteArray(num_actions, dtype=action_dtype),
        rewards=specs.Array((), reward_dtype),
        discounts=specs.BoundedArray((), reward_dtype, 0.0, 1.0)), **kwargs)


class ContinuousEnvironment(Environment):
  """`fakes.py:178-219`."""

  def __init__(self, *, action_dim: int = 1, observation_dim: int = 1, bounded: bool = False, dtype=np.float32,
               reward_dtype=np.float32, **kwargs):
    ashape = () if action_dim == 0 else (action_dim,)
    oshape = () if observation_dim == 0 else (observation_dim,)
    actions = specs.BoundedArray(ashape, dtype, -1.0, 1.0) if bounded else specs.Array(ashape, dtype)
    super().__init__(spec=specs.EnvironmentSpec(
        observations=specs.Array(oshape, dtype), actions=actions, rewards=specs.Array((), reward_dtype),
        discounts=specs.BoundedArray((), reward_dtype, 0.0, 1.0)), **kwargs)
