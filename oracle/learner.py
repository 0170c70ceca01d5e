"""Oracle: one learner update in PyTorch-CPU fp32 (TEST INFRASTRUCTURE, see oracle/__init__.py).

  DQNOracleLearner  <- acme/agents/tf/dqn/learning.py:112-168   
  D4PGOracleLearner <- acme/agents/tf/d4pg/learning.py:156-247  
  adam_update       <- snt.optimizers.Adam as recalled in SURVEY App. A.5 (paper form, eps_mode 0;
                       Keras "epsilon-hat" form as eps_mode 1).  Sonnet not in tree: UNPINNED.
The batch is passed in (already sampled / gathered); replay is a separate oracle.
"""

from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from oracle import losses as L


def bias_corrections(t: int, b1: float, b2: float):
  """1 - beta^t computed in f64 on the host, cast to f32 (shared definition with the kernel)."""
  return np.float32(1.0 - b1**t), np.float32(1.0 - b2**t)


class Adam:

  def __init__(self, lr, b1=0.9, b2=0.999, eps=1e-8, eps_mode=0):
    self.lr, self.b1, self.b2, self.eps, self.eps_mode = lr, b1, b2, eps, eps_mode
    self.t = 0
    self.m: Dict[str, np.ndarray] = {}
    self.v: Dict[str, np.ndarray] = {}

  def apply(self, grads: Dict[str, np.ndarray], params: Dict[str, np.ndarray]):
    f = np.float32
    self.t += 1
    bc1, bc2 = bias_corrections(self.t, self.b1, self.b2)
    out = {}
    for k, p in params.items():
      g = np.asarray(grads[k], f)
      m = self.m.get(k, np.zeros_like(p))
      v = self.v.get(k, np.zeros_like(p))
      m = f(self.b1) * m + f(1 - self.b1) * g
      v = f(self.b2) * v + f(1 - self.b2) * (g * g)
      # K7 flushes moments below FLT_MIN to zero (they would otherwise sit in the subnormal range for hundreds of
      # updates, where IEEE div/sqrt take a slow path on the GPU); restated here so both sides state one arithmetic.
      # Effect on the update: < 1e-30 absolute.
      tiny = f(np.finfo(np.float32).tiny)
      m = np.where(np.abs(m) < tiny, f(0), m)
      v = np.where(v < tiny, f(0), v)
      if self.eps_mode == 0:
        upd = (m / bc1) / (np.sqrt(v / bc2) + f(self.eps))
      else:
        upd = (np.sqrt(bc2) / bc1) * m / (np.sqrt(v) + f(self.eps))
      out[k] = (p - f(self.lr) * upd).astype(f)
      self.m[k], self.v[k] = m.astype(f), v.astype(f)
    return out


def preprocess_obs(o):
  """uint8 frames -> float32 / 255 (atari_wrapper.py:303-304 + single_precision.py); floats pass."""
  o = np.asarray(o)
  if o.dtype == np.uint8:
    return (o.astype(np.float32) / np.float32(255.0)).astype(np.float32)
  return o.astype(np.float32)


class DQNOracleLearner:

  def __init__(self, network, target_network, discount, importance_sampling_exponent,
               learning_rate, target_update_period, huber_loss_parameter=1.0,
               max_abs_reward=1.0, eps_mode=0, target_update_mode='pre_increment', is_weights_dtype='f64'):
    # target_update_mode: 'pre_increment' = TF learner (copy when num_steps % period == 0, tested BEFORE the
    # increment: dqn/learning.py:157-161); 'post_increment' = JAX learner (copy when (steps + 1) % period == 0,
    # jax/dqn/learning.py:114-119 + jax/utils.py:148-154).  is_weights_dtype: see losses.dqn_loss.
    assert target_update_mode in ('pre_increment', 'post_increment') and is_weights_dtype in ('f64', 'f32')
    self.target_update_mode, self.is_weights_dtype = target_update_mode, is_weights_dtype
    self.net, self.tgt = network, target_network
    self.discount = discount
    self.beta = importance_sampling_exponent
    self.delta = huber_loss_parameter
    self.period = target_update_period
    self.max_abs_reward = max_abs_reward
    self.opt = Adam(learning_rate, eps_mode=eps_mode)
    self.num_steps = 0

  def step(self, o_tm1, a_tm1, R, D, o_t, prob, global_wmax=None, grad_hook=None):
    x0 = torch.tensor(preprocess_obs(o_tm1))
    x1 = torch.tensor(preprocess_obs(o_t))
    q_tm1 = self.net(x0)
    with torch.no_grad():
      q_tv = self.tgt(x1)
      q_ts = self.net(x1)
    ref = L.dqn_loss(q_tm1.detach().numpy(), q_tv.numpy(), q_ts.numpy(), np.asarray(a_tm1, np.int64),
                     R, D, prob, self.discount, self.delta, self.beta, self.max_abs_reward,
                     global_wmax=global_wmax, weights_dtype=self.is_weights_dtype)
    # autograd through the same expression (independent check of dq_tm1)
    rows = torch.arange(q_tm1.shape[0])
    a = torch.tensor(np.asarray(a_tm1, np.int64))
    r = torch.clamp(torch.tensor(np.asarray(R, np.float32)), -self.max_abs_reward, self.max_abs_reward)
    d = torch.tensor(np.asarray(D, np.float32)) * np.float32(self.discount)
    best = torch.argmax(q_ts, dim=1)
    td = (r + d * q_tv[rows, best]) - q_tm1[rows, a]
    absx = td.abs()
    quad = torch.clamp(absx, max=self.delta)
    hub = 0.5 * quad * quad + self.delta * (absx - quad)
    loss = (hub * torch.tensor(ref['weight'])).mean()
    names = self.net.names()
    grads = torch.autograd.grad(loss, [self.net.vars[k] for k in names])
    gdict = {k: g.numpy() for k, g in zip(names, grads)}
    if grad_hook is not None:
      gdict = grad_hook(gdict)
    new = self.opt.apply(gdict, self.net.numpy())
    self.net.load(new)
    phase = 1 if self.target_update_mode == 'post_increment' else 0
    if (self.num_steps + phase) % self.period == 0:
      self.tgt.copy_from(self.net)
    self.num_steps += 1
    out = dict(ref)
    out.update(q_tm1=q_tm1.detach().numpy(), q_t_value=q_tv.numpy(), q_t_selector=q_ts.numpy(),
               loss_autograd=np.float32(loss.item()), grads=gdict)
    return out


class D4PGOracleLearner:

  def __init__(self, policy, critic, target_policy, target_critic, discount, target_update_period,
               policy_lr=1e-4, critic_lr=1e-4, clipping=True, eps_mode=0):
    self.policy, self.critic = policy, critic
    self.tpolicy, self.tcritic = target_policy, target_critic
    self.discount = np.float32(discount)
    self.period = target_update_period
    self.clipping = clipping
    self.popt = Adam(policy_lr, eps_mode=eps_mode)
    self.copt = Adam(critic_lr, eps_mode=eps_mode)
    self.num_steps = 0

  def _critic_loss(self, logits_tm1, logits_t, Rn, Dg):
    """losses.categorical (distributional.py:22-37) through autograd; `ref` carries the NumPy restatement's outputs."""
    ref = L.categorical(logits_tm1.detach().numpy(), logits_t.numpy(), self.critic.values.numpy(), Rn, Dg)
    target = torch.tensor(ref['target'])
    return ref, (-(target * torch.log_softmax(logits_tm1, dim=-1)).sum(dim=-1)).mean()

  def step(self, o_tm1, a_tm1, R, D, o_t):
    if self.num_steps % self.period == 0:
      self.tpolicy.copy_from(self.policy)
      self.tcritic.copy_from(self.critic)
    self.num_steps += 1
    o0 = torch.tensor(np.asarray(o_tm1, np.float32))
    a0 = torch.tensor(np.asarray(a_tm1, np.float32))
    o1 = torch.tensor(np.asarray(o_t, np.float32))
    Rn = np.asarray(R, np.float32)
    Dg = (self.discount * np.asarray(D, np.float32)).astype(np.float32)

    logits_tm1 = self.critic.logits(o0, a0)
    with torch.no_grad():
      logits_t = self.tcritic.logits(o1, self.tpolicy.action(o1))
    ref, critic_loss = self._critic_loss(logits_tm1, logits_t, Rn, Dg)

    a_t = self.policy.action(o1)
    a_leaf = a_t.detach().clone().requires_grad_(True)
    q = self.critic.mean(o1, a_leaf)
    dqda = torch.autograd.grad(q.sum(), a_leaf)[0].numpy()
    da, dqda_clipped = L.dpg_action_grad(dqda, 1.0 if self.clipping else None, self.clipping)
    policy_loss = np.float32(0.5 * (dqda_clipped**2).sum(axis=-1).mean())

    cn = self.critic.names()
    cg = torch.autograd.grad(critic_loss, [self.critic.vars[k] for k in cn])
    pn = self.policy.names()
    pg = torch.autograd.grad(a_t, [self.policy.vars[k] for k in pn], grad_outputs=torch.tensor(da))
    cg = {k: g.numpy() for k, g in zip(cn, cg)}
    pg = {k: g.numpy() for k, g in zip(pn, pg)}
    cnorm = pnorm = None
    if self.clipping:
      lst, pnorm = L.clip_by_global_norm([pg[k] for k in pn], 40.)
      pg = dict(zip(pn, lst))
      lst, cnorm = L.clip_by_global_norm([cg[k] for k in cn], 40.)
      cg = dict(zip(cn, lst))
    self.policy.load(self.popt.apply(pg, self.policy.numpy()))
    self.critic.load(self.copt.apply(cg, self.critic.numpy()))
    return dict(critic_loss=np.float32(critic_loss.item()), policy_loss=policy_loss,
                target=ref.get('target'), per_sample=ref['loss'], dlogits_tm1=ref.get('dlogits_tm1'), td=ref.get('td'),
                logits_tm1=logits_tm1.detach().numpy(), logits_t=logits_t.numpy(),
                dqda=dqda, critic_grads=cg, policy_grads=pg, critic_norm=cnorm, policy_norm=pnorm)


class DDPGOracleLearner(D4PGOracleLearner):
  """acme/agents/tf/ddpg/learning.py:140-237: D4PG's step with a scalar critic and trfl.td_learning (line 193)."""

  def _critic_loss(self, q_tm1, q_t, Rn, Dg):
    ref = L.td_learning(q_tm1.detach().numpy()[:, 0], Rn, Dg, q_t.numpy()[:, 0])
    target = torch.tensor(Rn) + torch.tensor(Dg) * q_t[:, 0]
    td = target - q_tm1[:, 0]
    return ref, (0.5 * td * td).mean()
