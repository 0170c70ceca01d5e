"""Oracle: prioritized replay table over a ring of steps (TEST INFRASTRUCTURE).

Reverb (dm-reverb-nightly==0.1.0.dev20200708, `setup.py:30`) is not in the reference tree:
parity UNPINNED.  Restated here are the table semantics Acme relies on
(`acme/agents/tf/dqn/agent.py:95-101`):
  Prioritized(alpha) sampler, Fifo remover at max_size, keys are opaque u64, updates to
  evicted / unknown keys are ignored, duplicates in one update call: last wins.

Storage model (this project's, mirrored by `acme_b200/csrc/replay_host.cpp`):
  * slot ring  (capacity S): one slot per *observation*; slot k of an episode holds
    (s_k, a_k, r_{k+1}, d_{k+1}) and a link to the slot of s_{k+1};
  * item ring  (capacity M = max_size): item = (start slot, end slot, length); its
    (R, D) are built at gather time by walking `length` links from the start slot with the
    arithmetic of `acme/adders/reverb/transition.py:135-145` (see oracle.nstep.nstep_return);
  * key = item sequence number; tree position = key mod M.
"""

from __future__ import annotations

import collections

import numpy as np

from oracle import sumtree
from oracle.nstep import nstep_return


class Writer:

  def __init__(self, table: 'Table'):
    self._t = table
    self.hist = collections.deque(maxlen=table.max_window + 1)  # slot seqs, newest last
    self.closed = False


class Table:

  def __init__(self, max_items, slot_capacity, obs_shape, obs_dtype, act_shape, act_dtype,
               gamma, alpha, max_window=8, shard_count=1):
    self.M = int(max_items)
    self.S = int(slot_capacity)
    self.max_window = max_window
    self.g = np.float32(gamma)
    self.alpha = float(alpha)
    self.shard_count = shard_count
    self.obs = np.zeros((self.S,) + tuple(obs_shape), obs_dtype)
    self.act = np.zeros((self.S,) + tuple(act_shape), act_dtype)
    self.rew = np.zeros(self.S, np.float32)
    self.disc = np.zeros(self.S, np.float32)
    self.next = np.full(self.S, -1, np.int32)
    self.item_start = np.zeros(self.M, np.int32)
    self.item_end = np.zeros(self.M, np.int32)
    self.item_len = np.zeros(self.M, np.int32)
    self.item_start_seq = np.zeros(self.M, np.int64)
    self.slot_head = 0
    self.item_head = 0
    self.item_tail = 0
    self.tree = sumtree.SumTree(self.M)

  # ---------------------------------------------------------------- insert path
  def writer(self) -> Writer:
    return Writer(self)

  def _alloc_slot(self, observation) -> int:
    seq = self.slot_head
    self.slot_head += 1
    idx = seq % self.S
    self.obs[idx] = observation
    self.next[idx] = -1
    floor = self.slot_head - self.S
    while self.item_tail < self.item_head and self.item_start_seq[self.item_tail % self.M] < floor:
      self.tree.set_leaves([self.item_tail % self.M], [0.0])
      self.item_tail += 1
    return seq

  def append(self, w: Writer, observation, action, reward, discount, next_observation):
    assert not w.closed
    if not w.hist:
      w.hist.append(self._alloc_slot(observation))
    cur = w.hist[-1] % self.S
    self.act[cur] = action
    self.rew[cur] = np.float32(reward)
    self.disc[cur] = np.float32(discount)
    nxt = self._alloc_slot(next_observation)
    self.next[cur] = nxt % self.S
    w.hist.append(nxt)

  def create_item(self, w: Writer, num_timesteps: int, priority: float) -> int:
    assert not w.closed
    if num_timesteps < 1 or len(w.hist) < num_timesteps + 1:
      raise ValueError('not enough timesteps appended')
    start_seq = w.hist[-(num_timesteps + 1)]
    end_seq = w.hist[-1]
    if start_seq < self.slot_head - self.S:
      raise ValueError('window already overwritten in the slot ring')
    key = self.item_head
    self.item_head += 1
    if self.item_head - self.item_tail > self.M:
      self.item_tail += 1
    pos = key % self.M
    self.item_start[pos] = start_seq % self.S
    self.item_end[pos] = end_seq % self.S
    self.item_len[pos] = num_timesteps
    self.item_start_seq[pos] = start_seq
    self.tree.set_leaves([pos], sumtree.weight_from_priority([priority], self.alpha))
    return key

  def close(self, w: Writer):
    w.closed = True
    w.hist.clear()

  # ---------------------------------------------------------------- sample path
  @property
  def size(self) -> int:
    return self.item_head - self.item_tail

  def sample(self, u, stratified=True):
    """-> keys u64[B], positions i64[B], probability f32[B] (leaf / (shards * mass))."""
    pos, prob = self.tree.sample(u, stratified)
    if self.shard_count != 1:
      prob = (self.tree.leaves[pos] / (np.float32(self.shard_count) * self.tree.total)).astype(np.float32)
    # key of the live item at a position: the unique seq in [tail, head) congruent to pos mod M
    base = self.item_tail - (self.item_tail % self.M)
    keys = base + pos
    keys = np.where(keys < self.item_tail, keys + self.M, keys).astype(np.uint64)
    return keys, pos, prob

  def gather(self, pos):
    pos = np.asarray(pos, np.int64)
    s = self.item_start[pos]
    e = self.item_end[pos]
    B = len(pos)
    R = np.zeros(B, np.float32)
    D = np.zeros(B, np.float32)
    for b in range(B):
      cur = int(s[b])
      rs, ds = [], []
      for _ in range(int(self.item_len[pos[b]])):
        rs.append(self.rew[cur])
        ds.append(self.disc[cur])
        cur = int(self.next[cur])
      R[b], D[b] = nstep_return(rs, ds, self.g)
    return self.obs[s], self.act[s], R, D, self.obs[e]

  def update_priorities(self, keys, priorities):
    keys = np.asarray(keys, np.uint64).astype(np.int64)
    w = sumtree.weight_from_priority(priorities, self.alpha)
    live = (keys >= self.item_tail) & (keys < self.item_head)
    self.tree.set_leaves(keys[live] % self.M, w[live])


class FrameStacker:
  """`acme/wrappers/frame_stacking.py:64-88` restated: the last `num_frames` frames stacked on a new LAST axis, newest
  last, blank (zero) frames before the first frame after `reset()`.  Used by the tests to produce the observation stream
  whose stacks the frame-deduplicated ring must rebuild bit for bit."""

  def __init__(self, num_frames: int):
    self.num_frames = num_frames
    self.reset()

  def reset(self):
    self._stack = []

  def step(self, frame):
    frame = np.asarray(frame)
    if not self._stack:
      self._stack = [np.zeros_like(frame)] * (self.num_frames - 1)
    self._stack = (self._stack + [frame])[-self.num_frames:]
    return np.stack(self._stack, axis=-1)
