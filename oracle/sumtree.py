"""Oracle: fan-out-32 fp32 sum-tree (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reverb's `Prioritized` selector is not in the reference tree (parity UNPINNED); what is
restated here is its contract as Acme uses it (`acme/agents/tf/dqn/agent.py:95-101`,
SURVEY App. A.3): stored weight = priority**alpha, P(i) = weight_i / sum(weights),
items with zero weight are never drawn, updates are applied in order (last wins).

The *arithmetic order* is this project's own and is shared bit-for-bit with the CUDA
kernels in `acme_b200/csrc/sumtree.cu`:
  * every node keeps the SEQUENTIAL fp32 inclusive prefixes of its 32 children,
    p_j = fl(p_{j-1} + c_j) (= np.cumsum on float32), and the node's own value is p_31;
  * stratified target  t_b = ((b + u_b) / B) * M, plain draw t_b = u_b * M  (fp32 ops);
  * descent: j = number of prefixes <= t (prefixes are monotone, so this is the first child
    with t < p_j, and such a child always has c_j > 0); if j == 32 (t >= node mass after
    rounding) j = last child whose prefix increased (0 if none); then t -= p_{j-1} (0 for j = 0);
  * probability = leaf / M (fp32 IEEE division).
"""

from __future__ import annotations

import numpy as np

F = 32
_f32 = np.float32


def num_levels(capacity: int) -> int:
  L = 1
  while F**L < capacity:
    L += 1
  return L


def level_width(capacity: int, L: int, lvl: int) -> int:
  """Number of entries stored at level `lvl` (1..L; L = leaves), padded to a multiple of 32."""
  span = F**(L - lvl)
  w = -(-capacity // span)
  return -(-w // F) * F


def seq_scan(x: np.ndarray) -> np.ndarray:
  """Sequential fp32 inclusive prefix along the last axis (length 32): p_j = fl(p_{j-1} + x_j)."""
  x = np.asarray(x, dtype=_f32)
  out = np.empty_like(x)
  acc = np.zeros(x.shape[:-1], _f32)
  for j in range(x.shape[-1]):
    acc = (acc + x[..., j]).astype(_f32)
    out[..., j] = acc
  return out


def weight_from_priority(p, alpha) -> np.ndarray:
  """Stored weight: float32(pow(float64(p), float64(alpha)))."""
  return np.power(np.asarray(p, dtype=np.float64), np.float64(alpha)).astype(_f32)


class SumTree:

  def __init__(self, capacity: int):
    self.capacity = int(capacity)
    self.L = num_levels(self.capacity)
    # levels[0] is the root (1 value); levels[l] for l=1..L
    self.levels = [np.zeros(1, _f32)] + [
        np.zeros(level_width(self.capacity, self.L, l), _f32) for l in range(1, self.L + 1)
    ]

  @property
  def leaves(self) -> np.ndarray:
    return self.levels[self.L]

  @property
  def total(self) -> np.float32:
    return self.levels[0][0]

  def rebuild(self):
    for l in range(self.L, 0, -1):
      child = self.levels[l].reshape(-1, F)
      sums = seq_scan(child)[:, F - 1]
      if l == 1:
        self.levels[0][0] = sums[0]
      else:
        self.levels[l - 1][:sums.shape[0]] = sums

  def set_leaves(self, positions, weights):
    """Sequential 'last wins' scatter + recompute of the touched ancestors."""
    positions = np.asarray(positions, dtype=np.int64)
    weights = np.asarray(weights, dtype=_f32)
    if positions.size == 0:
      return
    for p, w in zip(positions, weights):
      self.levels[self.L][p] = w
    touched = np.unique(positions)
    for l in range(self.L, 0, -1):
      touched = np.unique(touched // F)
      sums = seq_scan(self.levels[l].reshape(-1, F)[touched])[:, F - 1]
      if l == 1:
        self.levels[0][0] = sums[0]
      else:
        self.levels[l - 1][touched] = sums

  def targets(self, u: np.ndarray, stratified: bool) -> np.ndarray:
    u = np.asarray(u, dtype=_f32)
    B = u.shape[0]
    M = self.total
    if stratified:
      b = np.arange(B, dtype=_f32)
      return ((b + u) / _f32(B)) * M
    return u * M

  def sample(self, u: np.ndarray, stratified: bool = True):
    """Returns (leaf positions int64[B], probability f32[B])."""
    t = self.targets(u, stratified)
    B = t.shape[0]
    node = np.zeros(B, dtype=np.int64)
    rows = np.arange(B)
    for l in range(1, self.L + 1):
      child = self.levels[l].reshape(-1, F)[node]            # [B, 32]
      scan = seq_scan(child)
      j = (scan <= t[:, None]).sum(axis=1)                   # first child with t < prefix
      prev = np.concatenate([np.zeros((B, 1), _f32), scan[:, :-1]], axis=1)
      inc = scan > prev
      last_inc = np.where(inc.any(axis=1), (F - 1) - np.argmax(inc[:, ::-1], axis=1), 0)
      j = np.where(j == F, last_inc, j)
      excl = prev[rows, j]
      excl = np.where((j == 0), _f32(0), excl).astype(_f32)
      t = (t - excl).astype(_f32)
      node = node * F + j
    leaf = self.levels[self.L][node]
    with np.errstate(divide='ignore', invalid='ignore'):
      prob = (leaf / self.total).astype(_f32)
    return node, prob


class BinarySumTreeF64:
  """Reverb-like CPU baseline: binary tree, f64 nodes, one root-to-leaf walk per sample.

  Used only as the *throughput* baseline; not bit-compatible with `SumTree`.
  """

  def __init__(self, capacity: int):
    n = 1
    while n < capacity:
      n *= 2
    self.n = n
    self.tree = np.zeros(2 * n, np.float64)

  def set(self, idx, w):
    idx = np.asarray(idx, np.int64) + self.n
    w = np.asarray(w, np.float64)
    for i, v in zip(idx, w):
      self.tree[i] = v
      i //= 2
      while i >= 1:
        self.tree[i] = self.tree[2 * i] + self.tree[2 * i + 1]
        i //= 2

  def build(self, w):
    self.tree[self.n:self.n + len(w)] = w
    i = self.n
    while i > 1:
      half = i // 2
      self.tree[half:i] = self.tree[i:2 * i:2] + self.tree[i + 1:2 * i:2]
      i = half

  def sample(self, u):
    t = np.asarray(u, np.float64) * self.tree[1]
    out = np.empty(len(t), np.int64)
    for k, x in enumerate(t):
      i = 1
      while i < self.n:
        left = self.tree[2 * i]
        if x < left:
          i = 2 * i
        else:
          x -= left
          i = 2 * i + 1
      out[k] = i - self.n
    return out, self.tree[out + self.n] / self.tree[1]
