"""Host wiring without a GPU: EnvironmentLoop counts (acme/environment_loop_test.py:36-51), Agent
learner-step cadence (acme/agents/agent.py:45-89), Counter (acme/utils/counting_test.py)."""
import pytest

from acme_b200 import agent, counting, environment_loop, loggers, specs, testing

EPISODE_LENGTH = 10


def make_loop():
  env = testing.DiscreteEnvironment(episode_length=EPISODE_LENGTH)
  actor = testing.Actor(specs.make_environment_spec(env))
  return environment_loop.EnvironmentLoop(env, actor, logger=loggers.NoOpLogger()), actor


def test_one_episode():
  loop, _ = make_loop()
  result = loop.run_episode()
  assert result['episode_length'] == EPISODE_LENGTH and 'episode_return' in result and 'steps_per_second' in result


def test_run_episodes_and_steps():
  loop, actor = make_loop()
  loop.run(num_episodes=10)
  assert actor.num_updates == 10 * EPISODE_LENGTH
  loop, actor = make_loop()
  loop.run(num_steps=EPISODE_LENGTH + 5)
  assert actor.num_updates == 2 * EPISODE_LENGTH
  with pytest.raises(ValueError):
    loop.run(num_episodes=1, num_steps=1)


def test_episode_report_carries_counter_totals_and_call_order():
  """Per step the actor sees select_action -> observe -> update after one observe_first; the report merges the shared
  counter's running totals (environment_loop.py:68-104)."""
  env = testing.DiscreteEnvironment(episode_length=3)
  calls = []

  class Spy(testing.Actor):
    def select_action(self, observation):
      calls.append('select')
      return super().select_action(observation)
    def observe_first(self, timestep):
      calls.append('first')
      super().observe_first(timestep)
    def observe(self, action, next_timestep):
      calls.append('observe')
      super().observe(action, next_timestep)
    def update(self):
      calls.append('update')
      super().update()

  counter = counting.Counter()
  loop = environment_loop.EnvironmentLoop(env, Spy(specs.make_environment_spec(env)), counter=counter, logger=loggers.NoOpLogger())
  first = loop.run_episode()
  second = loop.run_episode()
  assert calls[:7] == ['first', 'select', 'observe', 'update', 'select', 'observe', 'update'] and calls.count('first') == 2
  assert (first['episodes'], first['steps'], second['episodes'], second['steps']) == (1, 3, 2, 6)
  assert second['episode_length'] == 3 and second['steps_per_second'] > 0
  assert counter.get_counts() == {'episodes': 2, 'steps': 6}


class CountingLearner:

  def __init__(self):
    self.steps = 0

  def step(self):
    self.steps += 1

  def get_variables(self, names):
    return []


@pytest.mark.parametrize('ops,min_obs,observations,expected', [
    (8.0, 1000, 1000, 1),      # DQN defaults: first update right after the 1000th observation (agent.py:161)
    (8.0, 1000, 1016, 3),      # ... then one every 8 observations
    (8.0, 1000, 999, 0),
    (0.25, 10, 12, 12),        # observations_per_step < 1 -> int(1/ops) learner steps per observation
])
def test_agent_cadence(ops, min_obs, observations, expected):
  env = testing.DiscreteEnvironment(episode_length=10**9)
  actor = testing.Actor(specs.make_environment_spec(env))
  learner = CountingLearner()
  a = agent.Agent(actor, learner, min_observations=min_obs, observations_per_step=ops)
  ts = env.reset()
  a.observe_first(ts)
  for _ in range(observations):
    action = a.select_action(ts.observation)
    ts = env.step(action)
    a.observe(action, ts)
    a.update()
  assert learner.steps == expected


def test_counter_hierarchy():
  parent = counting.Counter()
  child = counting.Counter(parent, prefix='learner', time_delta=0.)
  child.increment(steps=2)
  counts = child.increment(steps=3, walltime=1.5)
  assert counts['learner_steps'] == 5 and counts['learner_walltime'] == 1.5
  assert parent.get_counts()['learner_steps'] == 5
  state = parent.save()
  fresh = counting.Counter()
  fresh.restore(state)
  assert fresh.get_counts() == parent.get_counts()


# ---- acme/utils/counting_test.py:46-86 restated against acme_b200.counting
def test_counter_threading():
  import threading
  from acme_b200 import counting
  counter = counting.Counter()
  n = 10
  gate = threading.Barrier(n)

  def add():
    gate.wait()
    counter.increment(foo=1)
  threads = [threading.Thread(target=add) for _ in range(n)]
  for t in threads:
    t.start()
  for t in threads:
    t.join()
  assert counter.get_counts()['foo'] == n


def test_counter_caching():
  from acme_b200 import counting
  parent = counting.Counter()
  counter = counting.Counter(parent, time_delta=0.)
  counter.increment(foo=12)
  assert parent.get_counts() == counter.get_counts()


def test_counter_shared_counts():
  from acme_b200 import counting
  parent = counting.Counter()
  child1 = counting.Counter(parent, 'child1')
  child2 = counting.Counter(parent, 'child2')
  child1.increment(foo=1)
  assert child2.increment(foo=2) == {'child1_foo': 1, 'child2_foo': 2}


def test_ddpg_and_d4pg_agents_share_the_wiring_and_differ_in_the_learner(monkeypatch):
  """`ddpg/agent.py:36-173` and `d4pg/agent.py:36-180` build the same pieces (uniform table, n-step adder, dataset,
  noisy behaviour policy, cadence batch/samples_per_insert after max(batch, min_replay) observations) around their
  own learner.  Device objects are replaced by fakes: this checks the host wiring only."""
  from acme_b200 import d4pg, replay
  built = []

  class FakeLearner:
    def __init__(self, policy, critic, target_policy, target_critic, discount, period, dataset, **kwargs):
      built.append((type(self).__name__, policy, critic, target_policy, target_critic, discount, period, dataset, kwargs))
      self.steps = 0
    def step(self):
      self.steps += 1
    def get_variables(self, names):
      return [[] for _ in names]

  class FakeD4PGLearner(FakeLearner):
    pass

  class FakeDDPGLearner(FakeLearner):
    pass

  class Net:
    device = 0
    def clone(self):
      return Net()

  monkeypatch.setattr(d4pg.D4PG, '_learner_cls', FakeD4PGLearner)
  monkeypatch.setattr(d4pg.DDPG, '_learner_cls', FakeDDPGLearner)
  monkeypatch.setattr(replay, 'make_reverb_dataset', lambda **kw: ('dataset', kw))
  monkeypatch.setattr(d4pg, 'GaussianNoisePolicy', lambda policy, sigma, lo, hi, seed: (lambda observation: lo))
  assert d4pg.DDPG.__mro__[1] is d4pg.D4PG

  env = testing.ContinuousEnvironment(action_dim=2, observation_dim=3, bounded=True, episode_length=10)
  spec = specs.make_environment_spec(env)
  for cls, expected in ((d4pg.D4PG, 'FakeD4PGLearner'), (d4pg.DDPG, 'FakeDDPGLearner')):
    built.clear()
    policy, critic = Net(), Net()
    a = cls(spec, policy, critic, batch_size=4, min_replay_size=2, samples_per_insert=2.0, n_step=3, discount=0.9,
            target_update_period=7)
    name, p, c, tp, tc, discount, period, dataset, kwargs = built[0]
    assert name == expected and p is policy and c is critic and tp is not policy and tc is not critic
    assert (discount, period) == (0.9, 7) and dataset[0] == 'dataset' and dataset[1]['batch_size'] == 4
    assert dataset[1]['stratified'] is False            # Reverb's Uniform selector draws i.i.d.
    table = a._table
    assert table.alpha == 0.0 and table.max_window == 3 and table.min_size == 1
    assert a._schedule.period == 2 and a._schedule.burst == 1   # 4 / 2.0 observations per learner step
    a._server.stop()
