"""D4PG learner + agent on the B200 hot path.

`D4PGLearner.step()` restates `acme/agents/tf/d4pg/learning.py:156-247` as one stream of kernels:
  target copy (if due) -> K1 sample (uniform table) -> K3 gather/n-step -> critic(o_tm1, a_tm1),
  target_critic(o_t, target_policy(o_t)) -> K5 C51 projection + cross-entropy -> critic backward ->
  policy(o_t) -> critic(o_t, a) -> d mean / d a -> K(dpg) norm-clipped action gradient -> policy
  backward -> global-norm clip (40) of each gradient set -> two Adams.
The reference's D4PG has no importance weights and never writes priorities (uniform table,
`d4pg/agent.py:96-103`); neither does this one.
"""

from __future__ import annotations

import os
import time
from typing import List, Optional

import numpy as np

from acme_b200 import _capi, actors, adders, agent, core, counting, loggers, networks, parallel, replay, specs


class D4PGLearner(core.Learner, core.Saveable):

  def __init__(self, policy_network: networks.D4PGPolicy, critic_network: networks.D4PGCritic,
               target_policy_network: networks.D4PGPolicy, target_critic_network: networks.D4PGCritic,
               discount: float, target_update_period: int, dataset: replay.ReplayDataset,
               policy_lr: float = 1e-4, critic_lr: float = 1e-4, clipping: bool = True,
               counter: counting.Counter = None, logger: loggers.Logger = None, checkpoint: bool = True,
               eps_mode: int = 0, use_cuda_graph: bool = True, process_group=None):
    """`process_group`: data-parallel learner over per-rank replay shards -- both gradient sets are all-reduced (mean) over
    NCCL between the gradient half and the optimizer half of the step, then clipped and applied identically on every rank
    (the ordering of the reference's only multi-replica learner, `crr/recurrent_learning.py:346-359`)."""
    import torch
    self._torch = torch
    self._dp = parallel.DataParallel(process_group)
    self._policy, self._critic = policy_network, critic_network
    self._tpolicy, self._tcritic = target_policy_network, target_critic_network
    self._dataset = dataset
    self._discount = float(np.float32(discount))
    self._period = int(target_update_period)
    self._clipping = bool(clipping)
    self._plr, self._clr = float(np.float32(policy_lr)), float(np.float32(critic_lr))
    self._eps_mode = int(eps_mode)
    self._counter = counter or counting.Counter()
    self._logger = logger or loggers.TerminalLogger('learner', time_delta=1.)
    self._timestamp = None
    self._use_graph, self._graph, self._steps_done = bool(use_cuda_graph), None, 0

    dev = torch.device('cuda', critic_network.device)
    # the three independent chains of the step on three streams (see _gradient_half); B200RL_D4PG_STREAMS=0: one stream
    self._concurrent = os.environ.get('B200RL_D4PG_STREAMS', '1') != '0'
    self._side = [torch.cuda.Stream(device=dev) for _ in range(2)] if self._concurrent else None
    B = self.B = dataset.B
    K, A = critic_network.K, policy_network.act_dim
    f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    self._c_train, self._c_pi, self._c_tgt = (critic_network.make_buffers(B) for _ in range(3))
    self._p_online, self._p_tgt = policy_network.make_buffers(B), target_policy_network.make_buffers(B)
    self._cg_train, self._cg_pi = critic_network.make_grad_buffers(B), critic_network.make_grad_buffers(B)
    self._pg = policy_network.make_grad_buffers(B)
    self.target, self.dlogits, self.dlogits_pi = f32(B, K), f32(B, K), f32(B, K)
    self.critic_loss_ps, self.policy_loss_ps = f32(B), f32(B)
    self.critic_loss, self.policy_loss = f32(1), f32(1)
    self.dqda, self.da = f32(B, A), f32(B, A)
    self._pm, self._pv = torch.zeros_like(policy_network.params.flat), torch.zeros_like(policy_network.params.flat)
    self._cm, self._cv = torch.zeros_like(critic_network.params.flat), torch.zeros_like(critic_network.params.flat)
    self._num_steps = torch.zeros(1, dtype=torch.int64, device=dev)
    self._norm_ws, self._norm_ws2 = f32(1024), f32(1024)
    self._pscale, self._cscale, self.policy_norm, self.critic_norm = f32(1), f32(1), f32(1), f32(1)
    self._loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    self._loss_ring = torch.zeros(2, 2, dtype=torch.float32).pin_memory()      # fetch_loss='async': two slots
    self._loss_events = [torch.cuda.Event(), torch.cuda.Event()]
    self._loss_pending = None
    self._obs_dim, self._act_dim = critic_network.obs_dim, critic_network.act_dim

  def _views(self):
    torch, ds, B = self._torch, self._dataset, self.B
    o0 = ds.o_tm1.view(torch.float32).view(B, self._obs_dim)
    o1 = ds.o_t.view(torch.float32).view(B, self._obs_dim)
    a0 = ds.a_tm1.view(torch.float32).view(B, self._act_dim)
    return o0, a0, o1

  def _device_step_eager(self, uniforms=None):
    self._gradient_half(uniforms)
    self._exchange()
    self._apply_half()

  def _exchange(self):
    """Data parallel: mean of both gradient sets over the ranks (NCCL all-reduce; not capturable, sits between graphs)."""
    if not self._dp.enabled:
      return
    import torch.distributed as dist
    for net in (self._policy, self._critic):
      g = net.params.grad
      if dist.get_backend(self._dp.group) == 'nccl':
        dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self._dp.group)
      else:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self._dp.group)
        g.mul_(1.0 / self._dp.world)

  def _gradient_half(self, uniforms=None):
    st = _capi.current_stream()
    P, C, TP, TC = self._policy, self._critic, self._tpolicy, self._tcritic
    ds, B = self._dataset, self.B
    # learning.py:171-174: target <- online when num_steps % period == 0, before anything else
    for src, dst in ((P, TP), (C, TC)):
      _capi.call('b200rl_copy_if_period', src.params.size * 4, _capi.ptr(dst.params.flat), _capi.ptr(src.params.flat),
                 _capi.ptr(self._num_steps), self._period, 0, st)
    ds.sample_raw(uniforms)
    o0, a0, o1 = self._views()
    # The step is ~100 small launch-bound kernels; its three chains are independent until the critic loss, so they run
    # on three streams (fork / join by events, also inside the captured graph; one workspace lane per stream):
    #   main  critic(o_tm1, a_tm1)                         -> [join target] critic loss -> critic backward
    #   s1    target policy(o_t) -> target critic(o_t, .)
    #   s2    policy(o_t) -> critic(o_t, .) -> dq/da -> DPG loss -> policy backward
    torch = self._torch
    main = torch.cuda.current_stream()
    fork = self._concurrent
    if fork:
      s1, s2 = self._side
      ev = torch.cuda.Event()
      ev.record(main)
      s1.wait_event(ev)
      s2.wait_event(ev)
    else:
      s1 = s2 = main
    with torch.cuda.stream(s1):
      a_targ = TP.lane(1).action(o1, self._p_tgt)
      logits_t = TC.lane(1).logits(o1, a_targ, self._c_tgt)
      if fork:
        tgt_done = torch.cuda.Event()
        tgt_done.record(s1)
    with torch.cuda.stream(s2):
      # actor learning (learning.py:206-218): dq/da through the online critic, parameters untouched
      st2 = _capi.current_stream()
      a_t = P.lane(2).action(o1, self._p_online)
      logits_pi = C.lane(2).logits(o1, a_t, self._c_pi)
      self._dq_dlogits(logits_pi)
      C.backward_flat(self._c_pi['x'].data_ptr(), self._c_pi, self._cg_pi, self.dlogits_pi.data_ptr(),
                      param_grads=False, input_grad=True)
      _capi.call('b200rl_split_second', B, self._obs_dim, self._act_dim, _capi.ptr(self._cg_pi['dx']), _capi.ptr(self.dqda), st2)
      _capi.call('b200rl_dpg_action_grad', B, self._act_dim, _capi.ptr(self.dqda), 1.0 if self._clipping else 0.0,
                 int(self._clipping), 1.0 / B, _capi.ptr(self.da), _capi.ptr(self.policy_loss_ps), _capi.ptr(self.policy_loss), st2)
      P.backward_action(o1, self._p_online, self._pg, self.da)
      if fork:
        pi_done = torch.cuda.Event()
        pi_done.record(s2)
    # critic learning (learning.py:198-203)
    logits_tm1 = C.lane(0).logits(o0, a0, self._c_train)
    if fork:
      main.wait_event(tgt_done)
    self._critic_loss(logits_tm1, logits_t)
    C.backward_flat(self._c_train['x'].data_ptr(), self._c_train, self._cg_train, self.dlogits.data_ptr(),
                    param_grads=True, input_grad=False)
    if fork:
      main.wait_event(pi_done)
    for net in (P, C, TP, TC):
      net.lane(0)

  def _critic_loss(self, logits_tm1, logits_t):
    """K5: losses.categorical (distributional.py:22-37) and its gradient w.r.t. logits_tm1 (learning.py:202-203)."""
    C, ds, B = self._critic, self._dataset, self.B
    _capi.call('b200rl_c51_loss', B, C.K, C.vmin, C.vmax, _capi.ptr(logits_tm1), _capi.ptr(logits_t), _capi.ptr(ds.R),
               _capi.ptr(ds.D), self._discount, 1.0 / B, _capi.ptr(self.target), _capi.ptr(self.critic_loss_ps),
               _capi.ptr(self.dlogits), _capi.ptr(self.critic_loss), _capi.current_stream())

  def _dq_dlogits(self, logits_pi):
    """d mean(Z) / d logits of the critic's output at (o_t, policy(o_t)) (distributions.py:64-66; learning.py:206-208)."""
    C, B = self._critic, self.B
    _capi.call('b200rl_c51_mean_bwd', B, C.K, C.vmin, C.vmax, _capi.ptr(logits_pi), None, _capi.ptr(self.dlogits_pi),
               _capi.current_stream())

  def _apply_half(self):
    """clip each gradient set by its global norm (learning.py:235-237), then the two Adams (240-241), step counter."""
    torch = self._torch
    P, C = self._policy, self._critic
    main = torch.cuda.current_stream()

    def update(net, m, v, lr, ws, scale, norm):
      st = _capi.current_stream()
      gs = None
      if self._clipping:
        _capi.call('b200rl_global_norm_scale', net.params.size, _capi.ptr(net.params.grad), 40.0, _capi.ptr(ws),
                   _capi.ptr(scale), _capi.ptr(norm), st)
        gs = _capi.ptr(scale)
      _capi.call('b200rl_adam', net.params.size, _capi.ptr(net.params.flat), _capi.ptr(net.params.grad), _capi.ptr(m),
                 _capi.ptr(v), _capi.ptr(self._num_steps), lr, 0.9, 0.999, 1e-8, self._eps_mode, gs, None, st)

    # the two networks' clip + Adam are independent: policy on a side stream beside the critic
    if self._concurrent:
      side = self._side[0]
      ev = torch.cuda.Event()
      ev.record(main)
      side.wait_event(ev)
      with torch.cuda.stream(side):
        update(P, self._pm, self._pv, self._plr, self._norm_ws2, self._pscale, self.policy_norm)
        done = torch.cuda.Event()
        done.record(side)
      update(C, self._cm, self._cv, self._clr, self._norm_ws, self._cscale, self.critic_norm)
      main.wait_event(done)
    else:
      update(P, self._pm, self._pv, self._plr, self._norm_ws2, self._pscale, self.policy_norm)
      update(C, self._cm, self._cv, self._clr, self._norm_ws, self._cscale, self.critic_norm)
    _capi.call('b200rl_step_increment', _capi.ptr(self._num_steps), _capi.current_stream())

  def _device_step(self, uniforms=None):
    torch = self._torch
    if not self._use_graph or uniforms is not None or self._steps_done < 2:
      self._device_step_eager(uniforms)
      return
    if self._graph is None:
      def capture(fn):
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
          fn()
        return g
      if self._dp.enabled:     # the all-reduces sit between two graphs
        self._graph = (capture(lambda: self._gradient_half(None)), capture(self._apply_half))
      else:
        self._graph = (capture(lambda: self._device_step_eager(None)),)
    self._graph[0].replay()
    if self._dp.enabled:
      self._exchange()
      self._graph[1].replay()

  def step(self, uniforms=None, fetch_loss: bool = True):
    table = self._dataset.table
    table.flush()
    if table.size < 1:
      raise RuntimeError('replay is empty: MinSize(1) rate limiter would block')
    self._device_step(uniforms)
    self._steps_done += 1
    result = {}
    if fetch_loss == 'async':   # copy to pinned memory now, consume when the next step has been issued (or at drain())
      slot = self._steps_done & 1
      self._loss_ring[slot, 0:1].copy_(self.critic_loss, non_blocking=True)
      self._loss_ring[slot, 1:2].copy_(self.policy_loss, non_blocking=True)
      self._loss_events[slot].record()
      previous, self._loss_pending = self._loss_pending, slot
      if previous is not None:
        self._loss_events[previous].synchronize()
        result = {'critic_loss': float(self._loss_ring[previous, 0]), 'policy_loss': float(self._loss_ring[previous, 1])}
    elif fetch_loss:
      self._loss_host[0:1].copy_(self.critic_loss, non_blocking=True)
      self._loss_host[1:2].copy_(self.policy_loss, non_blocking=True)
      self._torch.cuda.current_stream().synchronize()
      result = {'critic_loss': float(self._loss_host[0]), 'policy_loss': float(self._loss_host[1])}
    timestamp = time.time()
    elapsed = timestamp - self._timestamp if self._timestamp else 0
    self._timestamp = timestamp
    result.update(self._counter.increment(steps=1, walltime=elapsed))
    self._logger.write(result)

  def drain(self):
    """The losses still in flight from a `fetch_loss='async'` step (or None)."""
    if self._loss_pending is None:
      return None
    slot, self._loss_pending = self._loss_pending, None
    self._loss_events[slot].synchronize()
    out = {'critic_loss': float(self._loss_ring[slot, 0]), 'policy_loss': float(self._loss_ring[slot, 1])}
    self._logger.write(out)
    return out

  def get_variables(self, names: List[str]) -> List[List[np.ndarray]]:
    nets = {'critic': self._tcritic, 'policy': self._tpolicy}   # learning.py:133-140: target networks
    return [list(nets[name].variables().values()) for name in names]

  @property
  def num_steps(self) -> int:
    return int(self._num_steps.item())

  def save(self):
    f = lambda t: t.cpu().numpy()
    return {'policy': f(self._policy.params.flat), 'critic': f(self._critic.params.flat),
            'target_policy': f(self._tpolicy.params.flat), 'target_critic': f(self._tcritic.params.flat),
            'policy_opt': (f(self._pm), f(self._pv)), 'critic_opt': (f(self._cm), f(self._cv)), 'num_steps': self.num_steps}

  def restore(self, state):
    t = self._torch.as_tensor
    self._policy.params.flat.copy_(t(state['policy']))
    self._critic.params.flat.copy_(t(state['critic']))
    self._tpolicy.params.flat.copy_(t(state['target_policy']))
    self._tcritic.params.flat.copy_(t(state['target_critic']))
    self._pm.copy_(t(state['policy_opt'][0])); self._pv.copy_(t(state['policy_opt'][1]))
    self._cm.copy_(t(state['critic_opt'][0])); self._cv.copy_(t(state['critic_opt'][1]))
    self._num_steps.fill_(int(state['num_steps']))


class DDPGLearner(D4PGLearner):
  """`acme/agents/tf/ddpg/learning.py:140-237`: the D4PG step with a scalar critic -- critic loss = trfl.td_learning
  (0.5 td^2, line 193) instead of the categorical projection, dq/da through the scalar output directly; same target-copy
  timing (before the update, 157-160), DPG loss with dq/da norm clipping, global-norm clip 40, two Adams."""

  def __init__(self, policy_network, critic_network, target_policy_network, target_critic_network, *args, **kwargs):
    super().__init__(policy_network, critic_network, target_policy_network, target_critic_network, *args, **kwargs)
    torch = self._torch
    dev = self.dlogits.device
    self.td = torch.zeros(self.B, dtype=torch.float32, device=dev)
    self._ones = torch.ones((self.B, 1), dtype=torch.float32, device=dev)

  def _critic_loss(self, q_tm1, q_t):
    ds, B = self._dataset, self.B
    _capi.call('b200rl_td_learning', B, _capi.ptr(q_tm1), _capi.ptr(q_t), _capi.ptr(ds.R), _capi.ptr(ds.D), self._discount,
               1.0 / B, _capi.ptr(self.td), _capi.ptr(self.critic_loss_ps), _capi.ptr(self.dlogits), _capi.ptr(self.critic_loss),
               _capi.current_stream())

  def _dq_dlogits(self, q_pi):
    self.dlogits_pi.copy_(self._ones)      # d q / d q = 1: the action gradient flows through the scalar output itself


class GaussianNoisePolicy:
  """policy -> ClippedGaussian(sigma) (`acme/agents/tf/d4pg/agent.py:117-123`, networks/noise.py:27-40)."""

  def __init__(self, policy: networks.D4PGPolicy, sigma: float, act_min, act_max, seed: int = 0):
    import torch
    self._p, self._sigma, self._torch = policy, float(sigma), torch
    self._lo, self._hi = np.asarray(act_min, np.float32), np.asarray(act_max, np.float32)
    self._rng = np.random.default_rng(seed)
    self._bufs = policy.make_buffers(1)

  def __call__(self, observation):
    obs = self._torch.as_tensor(np.ascontiguousarray(observation, dtype=np.float32).reshape(1, -1)).cuda(self._p.device)
    a = self._p.action(obs, self._bufs).cpu().numpy()[0]
    a = a + self._rng.standard_normal(a.shape).astype(np.float32) * self._sigma
    return np.clip(a, self._lo, self._hi).astype(np.float32)


class D4PG(agent.Agent):
  """`acme/agents/tf/d4pg/agent.py:36-180`: uniform table, n-step adder, dataset, actor, learner."""

  _learner_cls = D4PGLearner

  def __init__(self, environment_spec: specs.EnvironmentSpec, policy_network: networks.D4PGPolicy,
               critic_network: networks.D4PGCritic, discount: float = 0.99, batch_size: int = 256,
               prefetch_size: int = 4, target_update_period: int = 100, min_replay_size: int = 1000,
               max_replay_size: int = 1000000, samples_per_insert: float = 32.0, n_step: int = 5,
               sigma: float = 0.3, clipping: bool = True, logger: loggers.Logger = None, checkpoint: bool = False,
               seed: int = 0, use_cuda_graph: bool = True):
    table = replay.Table(name=replay.DEFAULT_PRIORITY_TABLE, sampler=replay.selectors.Uniform(),
                         remover=replay.selectors.Fifo(), max_size=max_replay_size,
                         rate_limiter=replay.rate_limiters.MinSize(1),
                         signature=adders.NStepTransitionAdder.signature(environment_spec), max_window=n_step,
                         discount=discount, device=critic_network.device)
    self._server = replay.Server([table], port=None)
    address = f'localhost:{self._server.port}'
    # the reference passes priority_fns={table: lambda x: 1.} (agent.py:108): the default does the same
    adder = adders.NStepTransitionAdder(client=replay.Client(address), n_step=n_step, discount=discount)
    dataset = replay.make_reverb_dataset(server_address=address, batch_size=batch_size, prefetch_size=prefetch_size,
                                         seed=seed, stratified=False)
    aspec = environment_spec.actions
    actor = actors.FeedForwardActor(GaussianNoisePolicy(policy_network, sigma, aspec.minimum, aspec.maximum, seed), adder)
    learner = self._learner_cls(policy_network, critic_network, policy_network.clone(), critic_network.clone(), discount,
                                target_update_period, dataset, clipping=clipping, logger=logger, checkpoint=checkpoint,
                                use_cuda_graph=use_cuda_graph)
    self._table = table
    super().__init__(actor=actor, learner=learner, min_observations=max(batch_size, min_replay_size),
                     observations_per_step=float(batch_size) / samples_per_insert)


class DDPG(D4PG):
  """`acme/agents/tf/ddpg/agent.py:36-173`: the same single-process wiring (uniform table, n-step adder with constant
  priorities, Gaussian-noise behaviour policy clipped to the action spec, two Adams at 1e-4) around `DDPGLearner`;
  `critic_network` is a scalar critic (`networks.DDPGCritic`)."""

  _learner_cls = DDPGLearner
