"""DQfD demonstration mixing on the GPU (`acme/agents/tf/dqfd/agent.py:111-122,160-219`) against
oracle.dqfd on the same draws: the replaced rows, their n-step reward / discount, keys and probabilities bit for bit; a
DQN learner fed by the mixed dataset against the oracle learner fed by the oracle's mixed batch; the agent in the loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def close(got, want, rtol=1e-5, atol_scale=1e-6, name=''):
  got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
  atol = atol_scale * max(float(np.abs(want).max()), 1e-30)
  np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=name)


def _episodes(rng, count, obs_shape, obs_dtype, num_actions, lengths=None, zero_discount_at=None):
  eps = []
  for e in range(count):
    L = int(lengths[e]) if lengths is not None else int(rng.integers(3, 30))
    if np.dtype(obs_dtype) == np.uint8:
      obs = rng.integers(0, 256, (L,) + tuple(obs_shape), dtype=np.uint8)
    else:
      obs = rng.standard_normal((L,) + tuple(obs_shape)).astype(obs_dtype)
    act = rng.integers(0, num_actions, L).astype(np.int32)
    rew = rng.standard_normal(L).astype(np.float32)
    disc = np.ones(L, np.float32)
    disc[-1] = 0.
    if zero_discount_at is not None and L > zero_discount_at + 1:
      disc[zero_discount_at] = 0.5          # an environment discount inside the window
    eps.append((obs, act, rew, disc))
  return eps


@pytest.mark.parametrize('n_step', [1, 3, 5])
def test_demo_mix_matches_oracle(n_step):
  import torch
  import helpers
  from acme_b200 import dqfd, replay
  from oracle import dqfd as odqfd
  rng = np.random.default_rng(10 + n_step)
  shape, A, B, ratio = (5,), 4, 96, 0.4
  spec, table, server, adder, oracle = helpers.make_pair(shape, np.float32, A, n_step, 0.99, 0.6, max_size=300)
  for ep in range(6):
    helpers.feed_episode(rng, adder, oracle, int(rng.integers(4, 20)), n_step, shape, np.float32, A)
  table.flush()
  helpers.sync_oracle_leaves(table, oracle)
  eps = _episodes(rng, 7, shape, np.float32, A, lengths=[3, 4, 29, 8, 3, 17, 6], zero_discount_at=2)
  ds = dqfd.MixedReplayDataset(table, B, eps, ratio, n_step=n_step, discount=0.99, seed=4)
  assert ds.demos.num_episodes == 7 and ds.demos.num_steps == 70
  u = rng.random(B).astype(np.float32)
  u3 = rng.random((B, 3)).astype(np.float32)
  # edges: the draw exactly at the ratio is NOT a demonstration; the last episode / the last admissible first step
  u3[0] = (np.float32(ratio), 0.5, 0.5)
  u3[1] = (np.nextafter(np.float32(ratio), np.float32(0)), 0.9999999, 0.9999999)
  u3[2] = (0., 0., 0.)
  u3[3] = (0.1, 0.0, 0.9999999)         # episode of length 3: only first = 0 exists
  u3[4] = (0.1, 2.5 / 7, 0.9999999)     # episode 2, first = max_index - 2: the window is cut at the episode's end
  ds.inject_demo_uniforms(torch.from_numpy(u3).cuda())
  ds.sample_raw(torch.from_numpy(u).cuda())
  torch.cuda.synchronize()
  keys, pos, prob = oracle.sample(u, True)
  o0, a, R, D, o1 = oracle.gather(pos)
  batch = dict(o_tm1=o0.copy(), a_tm1=np.asarray(a).copy(), R=R.copy(), D=D.copy(), o_t=o1.copy(),
               keys=np.asarray(keys, np.uint64).copy(), prob=np.asarray(prob, np.float32).copy())
  batch, is_demo = odqfd.mix(batch, eps, u3, ratio, n_step, 0.99)
  assert not is_demo[0] and is_demo[1:5].all() and 0.2 < is_demo.mean() < 0.6
  np.testing.assert_array_equal(ds.is_demo.cpu().numpy().astype(bool), is_demo)
  got = ds.as_sample()
  np.testing.assert_array_equal(got.data[0].cpu().numpy(), batch['o_tm1'])
  np.testing.assert_array_equal(got.data[4].cpu().numpy(), batch['o_t'])
  np.testing.assert_array_equal(got.data[1].cpu().numpy(), batch['a_tm1'])
  np.testing.assert_array_equal(got.data[2].cpu().numpy(), batch['R'])          # bit-exact: same fp32 operation order
  np.testing.assert_array_equal(got.data[3].cpu().numpy(), batch['D'])
  np.testing.assert_array_equal(ds.keys.cpu().numpy().view(np.uint64), batch['keys'])
  np.testing.assert_array_equal(ds.prob.cpu().numpy(), batch['prob'])
  assert (got.info.probability.cpu().numpy()[is_demo] == 1.0).all()
  server.stop()


@pytest.mark.parametrize('use_graph', [False, True])
def test_dqfd_learner_steps_match_oracle(use_graph):
  """The unchanged DQN learner on the mixed dataset (`dqfd/agent.py:141-149`): TD errors, losses, importance weights
  (demonstration rows enter with probability 1) and priorities against the oracle learner fed by oracle.dqfd.mix; the
  priority write-back must skip the demonstration rows (their key names no item)."""
  import torch
  import helpers
  from acme_b200 import _capi, dqfd, dqn, loggers, networks, replay
  from oracle import dqfd as odqfd
  from oracle import learner as olearner
  from oracle import nets as onets
  rng = np.random.default_rng(21)
  shape, A, n, B, ratio = (84, 84, 4), 6, 3, 16, 0.5
  spec, table, server, adder, oracle = helpers.make_pair(shape, np.uint8, A, n, 0.99, 0.6, max_size=300)
  for ep in range(8):
    helpers.feed_episode(rng, adder, oracle, int(rng.integers(5, 30)), n, shape, np.uint8, A)
  table.flush()
  helpers.sync_oracle_leaves(table, oracle)
  eps = _episodes(rng, 5, shape, np.uint8, A)
  net = networks.DQNAtariNetwork(A, seed=5)
  tgt = net.clone()
  onet, otgt = onets.DQNAtariNetwork(A), onets.DQNAtariNetwork(A)
  onet.load(net.variables())
  otgt.load(net.variables())
  ds = dqfd.MixedReplayDataset(table, B, eps, ratio, n_step=n, discount=0.99, seed=7)
  learner = dqn.DQNLearner(net, tgt, 0.99, 0.2, 1e-3, target_update_period=2, dataset=ds, replay_client=replay.Client(server),
                           logger=loggers.NoOpLogger(), use_cuda_graph=use_graph)
  ol = olearner.DQNOracleLearner(onet, otgt, 0.99, 0.2, 1e-3, 2)
  counter = torch.zeros(1, dtype=torch.int64, device='cuda')
  u_dev = torch.empty(B, device='cuda')
  u3_dev = torch.empty(3 * B, device='cuda')
  seen_demo = 0
  for step in range(4):
    _capi.call('b200rl_uniform', u_dev.data_ptr(), B, 7, counter.data_ptr(), step, _capi.current_stream())
    _capi.call('b200rl_uniform', u3_dev.data_ptr(), 3 * B, 7 ^ ds._SEED_SALT, counter.data_ptr(), step, _capi.current_stream())
    u, u3 = u_dev.cpu().numpy(), u3_dev.cpu().numpy().reshape(B, 3)
    keys, pos, prob = oracle.sample(u, True)
    o0, a, R, D, o1 = oracle.gather(pos)
    batch = dict(o_tm1=o0.copy(), a_tm1=np.asarray(a).copy(), R=R.copy(), D=D.copy(), o_t=o1.copy(),
                 keys=np.asarray(keys, np.uint64).copy(), prob=np.asarray(prob, np.float32).copy())
    batch, is_demo = odqfd.mix(batch, eps, u3, ratio, n, 0.99)
    seen_demo += int(is_demo.sum())
    ref = ol.step(batch['o_tm1'], batch['a_tm1'], batch['R'], batch['D'], batch['o_t'], batch['prob'])
    oracle.update_priorities(batch['keys'], ref['priority'])
    learner.step()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ds.is_demo.cpu().numpy().astype(bool), is_demo)
    np.testing.assert_array_equal(ds.keys.cpu().numpy().view(np.uint64), batch['keys'])
    np.testing.assert_array_equal(ds.R.cpu().numpy(), batch['R'])
    np.testing.assert_array_equal(ds.o_t.cpu().numpy().reshape((B,) + shape), batch['o_t'])
    close(learner.td.cpu().numpy(), ref['td'], atol_scale=2e-5, name=f'td step {step}')
    close(learner.priority.cpu().numpy(), ref['priority'], atol_scale=2e-5)
    close(learner.loss.cpu().numpy()[0], ref['loss'], rtol=1e-4)
    close(learner.weight.cpu().numpy(), ref['weight'])
    helpers.sync_oracle_leaves_loose(table, oracle)   # the trees agree: demonstration rows changed no priority
  assert 10 < seen_demo < 54 and learner.num_steps == 4
  server.stop()


def test_dqfd_agent_runs_in_the_environment_loop():
  """`acme/agents/tf/dqfd/agent_test.py`-shaped: the whole agent on a fake environment with fake demonstrations."""
  import torch
  from acme_b200 import dqfd, environment_loop, loggers, networks, specs, testing
  env = testing.DiscreteEnvironment(num_actions=5, num_observations=10, obs_dtype=np.float32, episode_length=10)
  spec = specs.make_environment_spec(env)
  rng = np.random.default_rng(3)
  demos = _episodes(rng, 4, (), np.float32, 5, lengths=[10, 10, 6, 3])
  net = networks.MLPQNetwork(1, [50, 50, spec.actions.num_values], seed=0)
  log = loggers.InMemoryLogger()
  agent = dqfd.DQfD(spec, net, demos, demonstration_ratio=0.5, batch_size=10, samples_per_insert=2, min_replay_size=10,
                    max_replay_size=1000, n_step=5, logger=log)
  loop = environment_loop.EnvironmentLoop(env, agent, logger=loggers.NoOpLogger())
  loop.run(num_episodes=4)
  torch.cuda.synchronize()
  assert agent._learner_obj.num_steps == 7
  assert len(log.data) == 7 and all(np.isfinite(d['loss']) for d in log.data)
  assert agent._table.size == 4 * 14
  # over 7 batches of 10 with ratio 0.5 some rows came from the demonstrations (nonzero rewards -> nonzero loss)
  assert any(d['loss'] > 0 for d in log.data)
