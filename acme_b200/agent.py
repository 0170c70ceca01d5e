"""Actor + Learner glue and learner-step cadence (`acme/agents/agent.py:28-92`)."""

from typing import List

from acme_b200 import core


class Agent(core.Actor, core.VariableSource):

  def __init__(self, actor: core.Actor, learner: core.Learner, min_observations: int,
               observations_per_step: float):
    self._actor, self._learner = actor, learner
    self._num_observations = -min_observations
    if observations_per_step >= 1.0:
      self._observations_per_update = int(observations_per_step)
      self._steps_per_update = 1
    else:
      self._observations_per_update = 1
      self._steps_per_update = int(1.0 / observations_per_step)

  def select_action(self, observation):
    return self._actor.select_action(observation)

  def observe_first(self, timestep):
    self._actor.observe_first(timestep)

  def observe(self, action, next_timestep):
    self._num_observations += 1
    self._actor.observe(action, next_timestep)

  def update(self):
    if self._num_observations >= 0 and self._num_observations % self._observations_per_update == 0:
      self._num_observations = 0
      for _ in range(self._steps_per_update):
        self._learner.step()
      self._actor.update()

  def get_variables(self, names: List[str]):
    return self._learner.get_variables(names)
