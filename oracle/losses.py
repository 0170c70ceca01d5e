"""Oracle: learner-side losses in NumPy (TEST INFRASTRUCTURE, see oracle/__init__.py).

  dqn_loss      <- acme/agents/tf/dqn/learning.py:127-154 + trfl.double_qlearning (public
                   formula; trfl not in tree -> UNPINNED) + acme/tf/losses/huber.py:48-57
  l2_project    <- acme/tf/losses/distributional.py:42-83 (dense [B,K,K] form, verbatim math)
  categorical   <- acme/tf/losses/distributional.py:22-37
  dpg_grad      <- acme/tf/losses/dpg.py:30-59 (gradient w.r.t. the action)
  clip_by_global_norm <- tf.clip_by_global_norm as used at acme/agents/tf/d4pg/learning.py:235-237
All fp32 unless the reference computes in f64 (importance weights, learning.py:138-140).
"""

from __future__ import annotations

import numpy as np

f32 = np.float32


def huber(x, delta):
  x = np.asarray(x, f32)
  absx = np.abs(x)
  quad = np.minimum(absx, f32(delta))
  lin = absx - quad
  return (f32(0.5) * quad * quad + f32(delta) * lin).astype(f32)


def dqn_loss(q_tm1, q_t_value, q_t_selector, a_tm1, R, D, prob, gamma, huber_delta=1.0,
             is_exponent=0.2, max_abs_reward=1.0, global_wmax=None, weights_dtype='f64'):
  """Returns dict(td, huber, weight, loss, priority, dq_tm1).

  weights_dtype: 'f64' = the TF learner (power and max in f64, then cast; dqn/learning.py:138-143);
  'f32' = the JAX learner (1/probs cast to f32 first, power and max in f32; jax/dqn/learning.py:94-96)."""
  q_tm1 = np.asarray(q_tm1, f32)
  q_t_value = np.asarray(q_t_value, f32)
  q_t_selector = np.asarray(q_t_selector, f32)
  B, A = q_tm1.shape
  rows = np.arange(B)
  r = np.clip(np.asarray(R, f32), f32(-max_abs_reward), f32(max_abs_reward))
  d = (np.asarray(D, f32) * f32(gamma)).astype(f32)
  best = np.argmax(q_t_selector, axis=1)            # first max wins, like tf.argmax
  target = (r + d * q_t_value[rows, best]).astype(f32)
  td = (target - q_tm1[rows, a_tm1]).astype(f32)
  h = huber(td, huber_delta)
  if weights_dtype == 'f32':
    w32 = (1.0 / np.asarray(prob, np.float64)).astype(f32)**f32(is_exponent)
    w64 = w32.astype(np.float64)      # reported maximum only
    wmax32 = np.max(w32) if global_wmax is None else f32(global_wmax)
    w = (w32 / wmax32).astype(f32)
  else:
    w64 = (1.0 / np.asarray(prob, np.float64))**np.float64(is_exponent)
    wmax = np.max(w64) if global_wmax is None else np.float64(global_wmax)
    w = (w64 / wmax).astype(f32)
  per_sample = (h * w).astype(f32)
  loss = per_sample.astype(np.float64).sum() / B      # reporting only; kernel sums fp32 in fixed order
  dq = np.zeros((B, A), f32)
  dq[rows, a_tm1] = -(w / f32(B)) * np.clip(td, f32(-huber_delta), f32(huber_delta))
  return dict(td=td, huber=h, weight=w, loss=f32(loss), per_sample=per_sample,
              priority=np.abs(td).astype(np.float64), dq_tm1=dq, wmax64=np.max(w64))


def softmax(x):
  x = np.asarray(x, f32)
  m = x.max(axis=-1, keepdims=True)
  e = np.exp(x - m)
  return (e / e.sum(axis=-1, keepdims=True)).astype(f32)


def l2_project(Zp, P, Zq):
  """Dense projection exactly as distributional.py:62-83 (including the concat edge trick)."""
  Zp = np.asarray(Zp, f32)
  P = np.asarray(P, f32)
  Zq = np.asarray(Zq, f32)
  vmin, vmax = Zq[0], Zq[-1]
  d_pos = np.concatenate([Zq, vmin[None]])[1:]
  d_neg = np.concatenate([vmax[None], Zq])[:-1]
  clipped_zp = np.clip(Zp, vmin, vmax)[:, None, :]
  clipped_zq = Zq[None, :, None]
  d_pos = (d_pos - Zq)[None, :, None]
  d_neg = (Zq - d_neg)[None, :, None]
  delta_qp = clipped_zp - clipped_zq
  d_sign = (delta_qp >= 0.).astype(f32)
  delta_hat = (d_sign * delta_qp / d_pos) - ((f32(1.) - d_sign) * delta_qp / d_neg)
  return np.sum(np.clip(f32(1.) - delta_hat, f32(0.), f32(1.)) * P[:, None, :], axis=2).astype(f32)


def categorical(logits_tm1, logits_t, values, R, Dg):
  """C51 loss: returns dict(target[B,K], loss[B], dlogits_tm1[B,K] of mean loss)."""
  values = np.asarray(values, f32)
  z = (np.asarray(R, f32)[:, None] + np.asarray(Dg, f32)[:, None] * values[None, :]).astype(f32)
  p = softmax(logits_t)
  target = l2_project(z, p, values)
  lt = np.asarray(logits_tm1, f32)
  m = lt.max(axis=-1, keepdims=True)
  lse = (m + np.log(np.exp(lt - m).sum(axis=-1, keepdims=True))).astype(f32)
  logp = lt - lse
  loss = -(target * logp).sum(axis=-1).astype(f32)
  B = lt.shape[0]
  dlogits = ((softmax(lt) * target.sum(axis=-1, keepdims=True) - target) / f32(B)).astype(f32)
  return dict(target=target, loss=loss, dlogits_tm1=dlogits)


def dpg_action_grad(dqda, clip=1.0, clip_norm=True):
  """dL/da for loss = mean_b 0.5*||sg(dqda+a)-a||^2 with per-sample norm clipping (dpg.py:41-57)."""
  dqda = np.asarray(dqda, f32)
  B = dqda.shape[0]
  if clip is not None:
    if clip_norm:
      # tf.clip_by_norm: t * clip / max(||t||_2, clip)
      nrm = np.sqrt((dqda * dqda).sum(axis=-1, keepdims=True)).astype(f32)
      dqda = ((dqda * f32(clip)) / np.maximum(nrm, f32(clip))).astype(f32)
    else:
      dqda = np.clip(dqda, -clip, clip)
  return (-dqda / f32(B)).astype(f32), dqda


def clip_by_global_norm(grads, clip):
  sq = sum(float((np.asarray(g, np.float64)**2).sum()) for g in grads)
  norm = np.sqrt(sq)
  scale = f32(clip / max(norm, clip))
  return [np.asarray(g, f32) * scale for g in grads], f32(norm)


def td_learning(v_tm1, r_t, pcont_t, v_t):
  """trfl.td_learning as called at acme/agents/tf/ddpg/learning.py:193 (public trfl formula; UNPINNED): target =
  r + pcont * v_t (stop-gradient), td = target - v_tm1, loss = 0.5 td^2.  Returns dict(td, loss[B], dv_tm1 of the mean)."""
  v_tm1, v_t = np.asarray(v_tm1, f32), np.asarray(v_t, f32)
  target = (np.asarray(r_t, f32) + np.asarray(pcont_t, f32) * v_t).astype(f32)
  td = (target - v_tm1).astype(f32)
  return dict(td=td, loss=(f32(0.5) * td * td).astype(f32), dv_tm1=(-td / f32(td.shape[0])).astype(f32))
