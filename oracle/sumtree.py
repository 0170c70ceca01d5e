"""Oracle: fan-out-32 fp32 sum-tree (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reverb's `Prioritized` selector is not in the reference tree (parity UNPINNED); what is
restated here is its contract as Acme uses it (`acme/agents/tf/dqn/agent.py:95-101`,
SURVEY App. A.3): stored weight = priority**alpha, P(i) = weight_i / sum(weights),
items with zero weight are never drawn, updates are applied in order (last wins).

The *arithmetic order* is this project's own and is shared bit-for-bit with the CUDA
kernels in `acme_b200/csrc/sumtree.cu`:
  * every internal node = element 31 of the Kogge-Stone inclusive scan of its 32 children
    (`x[i] += x[i-d]` for d = 1,2,4,8,16, fp32 round-to-nearest each add) -- this is what a
    warp computes with `__shfl_up_sync`;
  * stratified target  t_b = ((b + u_b) / B) * M, plain draw t_b = u_b * M  (fp32 ops);
  * descent: in a node pick the first child c with  t < scan[c]  and  child[c] > 0; if none,
    the last child with child[c] > 0; then t -= scan[c-1] (0 for c = 0);
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


def ks_scan(x: np.ndarray) -> np.ndarray:
  """Kogge-Stone inclusive scan along the last axis (length 32), fp32."""
  x = np.array(x, dtype=_f32, copy=True)
  for d in (1, 2, 4, 8, 16):
    x[..., d:] = x[..., d:] + x[..., :-d].copy()
  return x


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
      sums = ks_scan(child)[:, F - 1]
      if l == 1:
        self.levels[0][0] = sums[0]
      else:
        self.levels[l - 1][:sums.shape[0]] = sums

  def set_leaves(self, positions, weights):
    """Sequential 'last wins' scatter + recompute of the touched ancestors."""
    positions = np.asarray(positions, dtype=np.int64)
    weights = np.asarray(weights, dtype=_f32)
    for p, w in zip(positions, weights):
      self.levels[self.L][p] = w
    touched = np.unique(positions)
    for l in range(self.L, 0, -1):
      touched = np.unique(touched // F)
      sums = ks_scan(self.levels[l].reshape(-1, F)[touched])[:, F - 1]
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
      scan = ks_scan(child)
      ok = (t[:, None] < scan) & (child > 0)
      any_ok = ok.any(axis=1)
      first = np.argmax(ok, axis=1)
      nz = child > 0
      last_nz = np.where(nz.any(axis=1), (F - 1) - np.argmax(nz[:, ::-1], axis=1), 0)
      c = np.where(any_ok, first, last_nz)
      excl = np.where(c > 0, scan[rows, np.maximum(c - 1, 0)], _f32(0)).astype(_f32)
      t = (t - excl).astype(_f32)
      node = node * F + c
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
