"""End-to-end wiring on the GPU, shaped like the reference's agent smoke tests
(acme/agents/tf/dqn/agent_test.py:38-58, acme/agents/tf/d4pg/agent_test.py:62-83): build the whole
agent (table, adder, dataset, networks, learner) and run EnvironmentLoop on a fake environment."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_dqn_agent_runs_in_the_environment_loop():
  import torch
  from acme_b200 import dqn, environment_loop, loggers, networks, specs, testing
  env = testing.DiscreteEnvironment(num_actions=5, num_observations=10, obs_dtype=np.float32, episode_length=10)
  spec = specs.make_environment_spec(env)
  net = networks.MLPQNetwork(1, [50, 50, spec.actions.num_values], seed=0)
  log = loggers.InMemoryLogger()
  agent = dqn.DQN(spec, net, batch_size=10, samples_per_insert=2, min_replay_size=10, max_replay_size=1000, n_step=5,
                  logger=log)
  loop = environment_loop.EnvironmentLoop(env, agent, logger=loggers.NoOpLogger())
  loop.run(num_episodes=4)
  torch.cuda.synchronize()
  # 40 observations, first update after the 10th, then one every 5 (batch 10 / samples_per_insert 2)
  assert agent._learner_obj.num_steps == 7
  assert len(log.data) == 7 and all(np.isfinite(d['loss']) for d in log.data) and log.data[-1]['steps'] == 7
  # (the fake environment emits all-zero observations / rewards / discounts, so q == td == 0 and the parameters
  # legitimately do not move; like the reference test we only require that the whole loop runs)
  # 4 episodes x (10 + min(5,10) - 1) items
  assert agent._table.size == 4 * 14
  v = agent.get_variables(['policy'])
  assert len(v[0]) == 6 and v[0][0].shape == (1, 50)


def test_d4pg_agent_runs_in_the_environment_loop():
  import torch
  from acme_b200 import d4pg, environment_loop, loggers, networks, specs, testing
  env = testing.ContinuousEnvironment(episode_length=10, bounded=True)
  spec = specs.make_environment_spec(env)
  policy = networks.D4PGPolicy(1, 1, sizes=(10, 10), seed=0)
  critic = networks.D4PGCritic(1, 1, sizes=(10, 10), vmin=-150., vmax=150., num_atoms=51, seed=1)
  log = loggers.InMemoryLogger()
  agent = d4pg.D4PG(spec, policy, critic, batch_size=10, samples_per_insert=2, min_replay_size=10, max_replay_size=1000,
                    logger=log)
  loop = environment_loop.EnvironmentLoop(env, agent, logger=loggers.NoOpLogger())
  loop.run(num_episodes=3)
  torch.cuda.synchronize()
  assert len(log.data) == 5 and all(np.isfinite(d['critic_loss']) and np.isfinite(d['policy_loss']) for d in log.data)
  crit, pol = agent.get_variables(['critic', 'policy'])
  assert len(crit) == 8 and len(pol) == 8
