"""Protocol objects of the `dm_env` package that the hot-path seams mention.

The reference imports the real `dm_env` (`acme/core.py:26`, `acme/adders/base.py:21`).
It is not installable here, so this module re-provides the handful of names the
Adder / Actor / EnvironmentLoop seams use: `StepType`, `TimeStep`, the four
constructors, and the array specs.  If the real package is importable it is used
instead so that objects interoperate.
"""

from __future__ import annotations

import enum
from typing import Any, NamedTuple

import numpy as np

try:  # pragma: no cover - not available in this image
  from dm_env import (StepType, TimeStep, restart, transition, termination,  # type: ignore
                      truncation, Environment, specs)
  HAVE_DM_ENV = True
except Exception:  # noqa: BLE001
  HAVE_DM_ENV = False

  class StepType(enum.IntEnum):
    FIRST = 0
    MID = 1
    LAST = 2

    def first(self) -> bool:
      return self is StepType.FIRST

    def mid(self) -> bool:
      return self is StepType.MID

    def last(self) -> bool:
      return self is StepType.LAST

  class TimeStep(NamedTuple):
    step_type: Any
    reward: Any
    discount: Any
    observation: Any

    def first(self) -> bool:
      return self.step_type == StepType.FIRST

    def mid(self) -> bool:
      return self.step_type == StepType.MID

    def last(self) -> bool:
      return self.step_type == StepType.LAST

  def restart(observation):
    return TimeStep(StepType.FIRST, None, None, observation)

  def transition(reward, observation, discount=1.0):
    return TimeStep(StepType.MID, reward, discount, observation)

  def termination(reward, observation):
    return TimeStep(StepType.LAST, reward, 0.0, observation)

  def truncation(reward, observation, discount=1.0):
    return TimeStep(StepType.LAST, reward, discount, observation)

  class Environment:
    """Abstract environment: reset() / step(action) / *_spec()."""

    def reset(self) -> TimeStep:
      raise NotImplementedError

    def step(self, action) -> TimeStep:
      raise NotImplementedError

    def observation_spec(self):
      raise NotImplementedError

    def action_spec(self):
      raise NotImplementedError

    def reward_spec(self):
      return specs.Array(shape=(), dtype=float, name='reward')

    def discount_spec(self):
      return specs.BoundedArray(shape=(), dtype=float, minimum=0., maximum=1., name='discount')

    def close(self):
      pass

  class _Specs:
    """Namespace standing in for `dm_env.specs`."""

    class Array:

      def __init__(self, shape, dtype, name=None):
        self._shape = tuple(int(d) for d in shape)
        self._dtype = np.dtype(dtype)
        self._name = name

      shape = property(lambda self: self._shape)
      dtype = property(lambda self: self._dtype)
      name = property(lambda self: self._name)

      def __repr__(self):
        return f'Array(shape={self.shape}, dtype={self.dtype!r}, name={self.name!r})'

      def validate(self, value):
        value = np.asarray(value)
        if value.shape != self.shape:
          raise ValueError(f'Expected shape {self.shape} but found {value.shape}')
        if value.dtype != self.dtype:
          raise ValueError(f'Expected dtype {self.dtype} but found {value.dtype}')
        return value

      def generate_value(self):
        return np.zeros(shape=self.shape, dtype=self.dtype)

      def replace(self, **kwargs):
        args = dict(shape=self.shape, dtype=self.dtype, name=self.name)
        args.update(kwargs)
        return type(self)(**args)

    class BoundedArray(Array):

      def __init__(self, shape, dtype, minimum, maximum, name=None):
        super().__init__(shape, dtype, name)
        self._minimum = np.array(minimum, dtype=self.dtype)
        self._maximum = np.array(maximum, dtype=self.dtype)
        if np.any(self._minimum > self._maximum):
          raise ValueError('minimum > maximum')

      minimum = property(lambda self: self._minimum)
      maximum = property(lambda self: self._maximum)

      def validate(self, value):
        value = super().validate(value)
        if (value < self.minimum).any() or (value > self.maximum).any():
          raise ValueError('Values were not all within bounds')
        return value

      def generate_value(self):
        return (np.ones(shape=self.shape, dtype=self.dtype) * self.dtype.type(self.minimum))

      def replace(self, **kwargs):
        args = dict(shape=self.shape, dtype=self.dtype, minimum=self.minimum,
                    maximum=self.maximum, name=self.name)
        args.update(kwargs)
        return type(self)(**args)

    class DiscreteArray(BoundedArray):

      def __init__(self, num_values, dtype=np.int32, name=None):
        if num_values <= 0:
          raise ValueError('num_values must be positive')
        super().__init__(shape=(), dtype=dtype, minimum=0, maximum=num_values - 1, name=name)
        self._num_values = int(num_values)

      num_values = property(lambda self: self._num_values)

      def replace(self, **kwargs):
        args = dict(num_values=self.num_values, dtype=self.dtype, name=self.name)
        args.update(kwargs)
        return type(self)(**args)

  specs = _Specs()
