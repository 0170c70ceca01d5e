"""EnvironmentLoop: same wiring as `acme/environment_loop.py:29-144` (calls the Actor seam only)."""

import time
from typing import Optional

from acme_b200 import core, counting, loggers


class EnvironmentLoop(core.Worker):

  def __init__(self, environment, actor: core.Actor, counter: counting.Counter = None,
               logger: loggers.Logger = None, label: str = 'environment_loop'):
    self._environment = environment
    self._actor = actor
    self._counter = counter or counting.Counter()
    self._logger = logger or loggers.make_default_logger(label)

  def run_episode(self):
    start = time.time()
    steps, ret = 0, 0
    timestep = self._environment.reset()
    self._actor.observe_first(timestep)
    while not timestep.last():
      action = self._actor.select_action(timestep.observation)
      timestep = self._environment.step(action)
      self._actor.observe(action, next_timestep=timestep)
      self._actor.update()
      steps += 1
      ret += timestep.reward
    counts = self._counter.increment(episodes=1, steps=steps)
    result = {'episode_length': steps, 'episode_return': ret,
              'steps_per_second': steps / max(time.time() - start, 1e-9)}
    result.update(counts)
    return result

  def run(self, num_episodes: Optional[int] = None, num_steps: Optional[int] = None):
    if not (num_episodes is None or num_steps is None):
      raise ValueError('Either "num_episodes" or "num_steps" should be None.')
    episodes = steps = 0
    while not ((num_episodes is not None and episodes >= num_episodes) or
               (num_steps is not None and steps >= num_steps)):
      result = self.run_episode()
      episodes += 1
      steps += result['episode_length']
      self._logger.write(result)
