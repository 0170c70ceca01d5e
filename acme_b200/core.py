"""Actor / Learner / Saveable interfaces kept from the reference (`acme/core.py:31-152`).

Only the abstract seams are restated; everything that implements them in this
package is B200-native.
"""

import abc
from typing import Generic, List, TypeVar

T = TypeVar('T')


class Actor(abc.ABC):
  """select_action / observe_first / observe / update (`acme/core.py:31-81`)."""

  @abc.abstractmethod
  def select_action(self, observation):
    ...

  @abc.abstractmethod
  def observe_first(self, timestep):
    ...

  @abc.abstractmethod
  def observe(self, action, next_timestep):
    ...

  @abc.abstractmethod
  def update(self):
    ...


class VariableSource(abc.ABC):
  """get_variables(names) -> list of numpy nests (`acme/core.py:87-106`)."""

  @abc.abstractmethod
  def get_variables(self, names: List[str]):
    ...


class Worker(abc.ABC):

  @abc.abstractmethod
  def run(self):
    ...


class Learner(VariableSource, Worker):
  """One SGD update per `step()`; `run()` loops forever (`acme/core.py:117-140`)."""

  @abc.abstractmethod
  def step(self):
    ...

  def run(self):
    while True:
      self.step()


class Saveable(abc.ABC, Generic[T]):
  """save() / restore(state) (`acme/core.py:143-152`)."""

  @abc.abstractmethod
  def save(self) -> T:
    ...

  @abc.abstractmethod
  def restore(self, state: T):
    ...
