"""FeedForwardActor (`acme/agents/tf/actors.py:35-94`): batch-1 policy forward on the learner's own
network object (actor and learner share parameters, like the single-process reference agent),
forwarding observations to the adder."""

from typing import Callable, Optional

import numpy as np

from acme_b200 import core


class FeedForwardActor(core.Actor):

  def __init__(self, policy: Callable[[np.ndarray], np.ndarray], adder=None, variable_client=None):
    self._policy, self._adder, self._variable_client = policy, adder, variable_client

  def select_action(self, observation):
    return self._policy(observation)

  def observe_first(self, timestep):
    if self._adder:
      self._adder.add_first(timestep)

  def observe(self, action, next_timestep):
    if self._adder:
      self._adder.add(action, next_timestep)

  def update(self):
    if self._variable_client:
      self._variable_client.update()


class EpsilonGreedyPolicy:
  """snt.Sequential([network, lambda q: trfl.epsilon_greedy(q, eps).sample()])
  (`acme/agents/tf/dqn/agent.py:119-124`): greedy w.p. 1-eps, uniform otherwise."""

  def __init__(self, network, epsilon: float = 0.05, seed: int = 0):
    import torch
    self._net, self._eps = network, float(epsilon)
    self._rng = np.random.default_rng(seed)
    self._bufs = network.make_buffers(1)
    self._torch = torch

  def __call__(self, observation):
    torch = self._torch
    if self._rng.random() < self._eps:
      return np.int32(self._rng.integers(self._net.A))
    obs = torch.as_tensor(np.ascontiguousarray(observation)[None]).cuda(self._net.device)
    q = self._net.forward(obs, self._bufs)
    return np.int32(int(np.argmax(q.cpu().numpy()[0])))  # first max wins, like tf.argmax
