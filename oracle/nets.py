"""Oracle: the Q-network / critic / policy modules in PyTorch-CPU fp32 (TEST INFRASTRUCTURE).

Restates (Sonnet/TF are not installable; layer semantics from their public behaviour, parity
UNPINNED — SURVEY §8c):
  AtariTorso / DQNAtariNetwork   acme/tf/networks/atari.py:36-69   (snt.Conv2D default padding SAME)
  DuellingMLP                    acme/tf/networks/duelling.py:27-59
  bsuite MLP                     examples/bsuite/run_dqn.py:46-49   (Flatten, MLP[50,50,A])
  LayerNormMLP                   acme/tf/networks/continuous.py:37-68
  CriticMultiplexer              acme/tf/networks/multiplexers.py:58-80 (concat obs, act)
  DiscreteValuedHead / mean      acme/tf/networks/distributional.py:58-67, distributions.py:64-66
  TanhToSpec                     acme/tf/networks/rescaling.py:63-74

Variables are exchanged as {name: ndarray} in *Sonnet shapes* (conv HWIO, linear [in,out]).
"""

from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as Fnn


def tf_same_pad(size, k, s):
  out = -(-size // s)
  total = max((out - 1) * s + k - size, 0)
  return out, total // 2, total - total // 2


def truncated_normal(rng: np.random.Generator, shape, std):
  x = rng.standard_normal(shape)
  bad = np.abs(x) > 2
  while bad.any():
    x[bad] = rng.standard_normal(int(bad.sum()))
    bad = np.abs(x) > 2
  return (x * std).astype(np.float32)


class _Net:
  """Holds named fp32 leaf tensors (requires_grad) in Sonnet shapes."""

  def __init__(self):
    self.vars: Dict[str, torch.Tensor] = {}

  def load(self, variables: Dict[str, np.ndarray]):
    assert set(variables) == set(self.vars), (sorted(variables), sorted(self.vars))
    for k, v in variables.items():
      assert tuple(v.shape) == tuple(self.vars[k].shape), (k, v.shape, self.vars[k].shape)
      self.vars[k] = torch.tensor(np.asarray(v, np.float32), requires_grad=True)

  def numpy(self) -> Dict[str, np.ndarray]:
    return {k: v.detach().numpy().copy() for k, v in self.vars.items()}

  def names(self):
    return list(self.vars)

  def _add(self, name, arr):
    self.vars[name] = torch.tensor(arr, requires_grad=True)

  def copy_from(self, other: '_Net'):
    for k in self.vars:
      self.vars[k] = other.vars[k].detach().clone().requires_grad_(True)


def _conv_same(x_nhwc, w_hwio, b, stride):
  # x: [B,H,W,C] -> NCHW conv with explicit asymmetric TF-SAME padding -> NHWC
  kh, kw = w_hwio.shape[0], w_hwio.shape[1]
  _, pt, pb = tf_same_pad(x_nhwc.shape[1], kh, stride)
  _, pl, pr = tf_same_pad(x_nhwc.shape[2], kw, stride)
  x = x_nhwc.permute(0, 3, 1, 2)
  x = Fnn.pad(x, (pl, pr, pt, pb))
  y = Fnn.conv2d(x, w_hwio.permute(3, 2, 0, 1), b, stride=stride)
  return y.permute(0, 2, 3, 1)


class DQNAtariNetwork(_Net):

  def __init__(self, num_actions, seed=0, in_hw=84, in_c=4):
    super().__init__()
    rng = np.random.default_rng(seed)
    self.A = num_actions
    h = in_hw
    geo = [(8, 4, in_c, 32), (4, 2, 32, 64), (3, 1, 64, 64)]
    for i, (k, s, ci, co) in enumerate(geo, 1):
      self._add(f'conv{i}/w', truncated_normal(rng, (k, k, ci, co), 1 / math.sqrt(k * k * ci)))
      self._add(f'conv{i}/b', np.zeros(co, np.float32))
      h = tf_same_pad(h, k, s)[0]
    self.flat = h * h * 64
    for head, n in (('value', 1), ('adv', num_actions)):
      self._add(f'{head}/l0/w', truncated_normal(rng, (self.flat, 512), 1 / math.sqrt(self.flat)))
      self._add(f'{head}/l0/b', np.zeros(512, np.float32))
      self._add(f'{head}/l1/w', truncated_normal(rng, (512, n), 1 / math.sqrt(512)))
      self._add(f'{head}/l1/b', np.zeros(n, np.float32))

  def __call__(self, obs_nhwc_f32: torch.Tensor) -> torch.Tensor:
    v = self.vars
    x = obs_nhwc_f32
    for i, s in ((1, 4), (2, 2), (3, 1)):
      x = torch.relu(_conv_same(x, v[f'conv{i}/w'], v[f'conv{i}/b'], s))
    x = x.reshape(x.shape[0], -1)                     # (h, w, c) order
    val = torch.relu(x @ v['value/l0/w'] + v['value/l0/b']) @ v['value/l1/w'] + v['value/l1/b']
    adv = torch.relu(x @ v['adv/l0/w'] + v['adv/l0/b']) @ v['adv/l1/w'] + v['adv/l1/b']
    adv = adv - adv.mean(dim=-1, keepdim=True)
    return val + adv


class MLPQNetwork(_Net):
  """snt.Sequential([Flatten, MLP(sizes)]) with ReLU between layers."""

  def __init__(self, in_dim, sizes, seed=0):
    super().__init__()
    rng = np.random.default_rng(seed)
    d = in_dim
    self.n = len(sizes)
    for i, h in enumerate(sizes):
      self._add(f'l{i}/w', truncated_normal(rng, (d, h), 1 / math.sqrt(d)))
      self._add(f'l{i}/b', np.zeros(h, np.float32))
      d = h

  def __call__(self, obs: torch.Tensor) -> torch.Tensor:
    x = obs.reshape(obs.shape[0], -1)
    for i in range(self.n):
      x = x @ self.vars[f'l{i}/w'] + self.vars[f'l{i}/b']
      if i + 1 < self.n:
        x = torch.relu(x)
    return x


def _uniform_fan_out(rng, shape, scale=0.333):
  # tf VarianceScaling(scale, mode='fan_out', distribution='uniform'): limit = sqrt(3*scale/fan_out)
  lim = math.sqrt(3.0 * scale / shape[1])
  return rng.uniform(-lim, lim, shape).astype(np.float32)


class LayerNormMLP(_Net):
  """Linear -> LayerNorm(eps=1e-5, scale+offset) -> tanh -> [Linear -> ELU]* (+ optional head)."""

  def __init__(self, in_dim, sizes, head_dim=None, head_name='head', head_scale=None,
               activate_final=True, seed=0):
    super().__init__()
    rng = np.random.default_rng(seed)
    self.sizes = list(sizes)
    self.head_name = head_name if head_dim else None
    self.activate_final = activate_final
    d = in_dim
    for i, h in enumerate(sizes):
      self._add(f'l{i}/w', _uniform_fan_out(rng, (d, h)))
      self._add(f'l{i}/b', np.zeros(h, np.float32))
      if i == 0:
        self._add('ln/scale', np.ones(h, np.float32))
        self._add('ln/offset', np.zeros(h, np.float32))
      d = h
    if head_dim:
      if head_scale is None:
        w = truncated_normal(rng, (d, head_dim), 1 / math.sqrt(d))
      else:  # NearZeroInitializedLinear: VarianceScaling(scale) -> truncated normal, fan_in
        w = truncated_normal(rng, (d, head_dim), math.sqrt(head_scale / d) / .87962566103423978)
      self._add(f'{head_name}/w', w)
      self._add(f'{head_name}/b', np.zeros(head_dim, np.float32))

  def torso(self, x):
    v = self.vars
    x = x @ v['l0/w'] + v['l0/b']
    x = Fnn.layer_norm(x, (x.shape[-1],), v['ln/scale'], v['ln/offset'], eps=1e-5)
    x = torch.tanh(x)
    for i in range(1, len(self.sizes)):
      x = x @ v[f'l{i}/w'] + v[f'l{i}/b']
      if i + 1 < len(self.sizes) or self.activate_final:
        x = Fnn.elu(x)
    return x

  def __call__(self, x):
    x = self.torso(x)
    if self.head_name:
      x = x @ self.vars[f'{self.head_name}/w'] + self.vars[f'{self.head_name}/b']
    return x


class D4PGCritic(LayerNormMLP):
  """CriticMultiplexer -> LayerNormMLP(sizes, activate_final) -> DiscreteValuedHead(K)."""

  def __init__(self, obs_dim, act_dim, sizes=(512, 512, 256), vmin=-150., vmax=150., atoms=51, seed=0):
    super().__init__(obs_dim + act_dim, sizes, head_dim=atoms, head_name='head', seed=seed)
    self.values = torch.tensor(np.linspace(vmin, vmax, atoms, dtype=np.float32))

  def logits(self, obs, act):
    return self(torch.cat([obs.reshape(obs.shape[0], -1), act.reshape(act.shape[0], -1)], dim=-1))

  def mean(self, obs, act):
    return (torch.softmax(self.logits(obs, act), dim=-1) * self.values).sum(dim=-1)


class DDPGCritic(LayerNormMLP):
  """CriticMultiplexer(critic_network=LayerNormMLP(sizes + [1])) (acme/agents/tf/ddpg/agent_test.py:46-47): the hidden
  sizes are activated (ELU), the last layer of width 1 is not -- the same as a torso with `activate_final` plus a linear
  output."""

  def __init__(self, obs_dim, act_dim, sizes=(512, 512, 256), seed=0):
    super().__init__(obs_dim + act_dim, sizes, head_dim=1, head_name='q', seed=seed)

  def logits(self, obs, act):
    return self(torch.cat([obs.reshape(obs.shape[0], -1), act.reshape(act.shape[0], -1)], dim=-1))

  def mean(self, obs, act):
    return self.logits(obs, act)[:, 0]


class D4PGPolicy(LayerNormMLP):
  """LayerNormMLP(sizes, activate_final) -> NearZeroInitializedLinear(A) -> TanhToSpec."""

  def __init__(self, obs_dim, act_dim, sizes=(256, 256, 256), act_min=-1., act_max=1., seed=0):
    super().__init__(obs_dim, sizes, head_dim=act_dim, head_name='out', head_scale=1e-4, seed=seed)
    self.scale = torch.tensor(np.broadcast_to(np.float32(act_max) - np.float32(act_min), (act_dim,)).copy())
    self.offset = torch.tensor(np.broadcast_to(np.float32(act_min), (act_dim,)).copy())

  def action(self, obs):
    x = torch.tanh(self(obs.reshape(obs.shape[0], -1)))
    x = 0.5 * (x + 1.0)
    return x * self.scale + self.offset
