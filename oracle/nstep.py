"""Oracle: n-step transition assembly (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, independently written, what the reference's actor-side adder does:
  * episode state machine        -> acme/adders/reverb/base.py:126-176
  * n-step return / discount     -> acme/adders/reverb/transition.py:119-165
  * end-of-episode drain         -> acme/adders/reverb/transition.py:167-172
  * recording fake client/writer -> acme/adders/reverb/test_utils.py:32-74

Two forms are given:
  1. `ReferenceAdder` + `RecordingClient`: object form, drives the same
     add_first/add protocol and records materialised `(o, a, R, D, o')` items,
     used against the reference's 7 golden cases.
  2. `enumerate_items` / `nstep_return`: array form of SURVEY App. A.1 / A.2
     (which (start, length) windows an episode produces, and the fp32 arithmetic
     of one window), used against the CUDA gather kernel.
"""

from __future__ import annotations

import collections
from typing import List, Sequence, Tuple

import numpy as np

DEFAULT_TABLE = 'priority_table'  # acme/adders/reverb/base.py:30


# ----------------------------------------------------------------------------- array form
def nstep_return(rewards: Sequence, discounts: Sequence, g) -> Tuple[np.ndarray, np.ndarray]:
  """(R, D) of one window, arithmetic order of transition.py:135-145.

  R = r_0; D = d_0; for j>=1: D*=g; R+=r_j*D; D*=d_j   -- each op rounded on its own
  (NumPy scalar arithmetic never fuses multiply-add).  dtype follows the inputs
  (fp32 inputs stay fp32 because `g` is np.float32, transition.py:111).
  """
  g = np.float32(g)
  R = np.array(rewards[0]).copy()
  D = np.array(discounts[0]).copy()
  for j in range(1, len(rewards)):
    D = D * g
    R = R + rewards[j] * D
    D = D * discounts[j]
  return R, D


def enumerate_items(T: int, n: int) -> List[Tuple[int, int]]:
  """(start, length) of every item one episode of T steps yields, in insert order.

  App. A.1: after the k-th `add` the window is the whole deque -> (max(0,k-n), min(k,n));
  when step T is terminal the deque is drained oldest-first -> (T-m+j, m-j), j=1..m-1,
  m=min(n,T).  Total T + m - 1 items.
  """
  items = [(max(0, k - n), min(k, n)) for k in range(1, T + 1)]
  m = min(n, T)
  items += [(T - m + j, m - j) for j in range(1, m)]
  return items


# ----------------------------------------------------------------------------- object form
class RecordingWriter:
  """Stands in for `reverb.Writer`: remembers what was appended / itemised."""

  def __init__(self, max_sequence_length, delta_encoded=False, chunk_length=None):
    self.max_sequence_length = max_sequence_length
    self.delta_encoded = delta_encoded
    self.chunk_length = chunk_length
    self.timesteps = []
    self.priorities = []  # (table, item, priority)
    self.closed = False

  def append(self, timestep):
    assert not self.closed
    self.timesteps.append(timestep)

  def create_item(self, table, num_timesteps, priority):
    assert not self.closed
    assert num_timesteps <= len(self.timesteps)
    assert num_timesteps <= self.max_sequence_length
    tail = self.timesteps[-num_timesteps:]
    self.priorities.append((table, tail[0] if num_timesteps == 1 else tail, priority))

  def close(self):
    assert not self.closed
    self.closed = True


class RecordingClient:
  """Stands in for `reverb.Client`: hands out `RecordingWriter`s and keeps them."""

  def __init__(self):
    self.writers: List[RecordingWriter] = []

  def writer(self, max_sequence_length, delta_encoded=False, chunk_length=None):
    w = RecordingWriter(max_sequence_length, delta_encoded, chunk_length)
    self.writers.append(w)
    return w


_Pending = collections.namedtuple('_Pending', 'observation action reward discount extras')


class ReferenceAdder:
  """CPU restatement of `NStepTransitionAdder` writing materialised transitions."""

  def __init__(self, client, n_step: int, discount: float, priority_fns=None):
    self._client = client
    self._g = np.float32(discount)                      # transition.py:111
    self._window = collections.deque(maxlen=n_step)     # base.py:107 (buffer_size=n_step)
    self._dangling = None                               # the not-yet-acted-on observation
    self._writer_obj = None
    self._priority_fns = dict(priority_fns) if priority_fns else {DEFAULT_TABLE: lambda x: 1.}

  # writer is created on first use and dropped by reset()  (base.py:111-132)
  def _writer(self):
    if self._writer_obj is None:
      self._writer_obj = self._client.writer(1, delta_encoded=False, chunk_length=None)
    return self._writer_obj

  def reset(self):
    if self._writer_obj is not None:
      self._writer_obj.close()
      self._writer_obj = None
    self._window.clear()
    self._dangling = None

  def add_first(self, timestep):
    if not timestep.first():
      raise ValueError('add_first needs a FIRST timestep')
    if self._dangling is not None:
      raise ValueError('reset must precede add_first')
    self._dangling = timestep.observation

  def add(self, action, next_timestep, extras=()):
    if self._dangling is None:
      raise ValueError('add_first must precede add')
    self._window.append(_Pending(self._dangling, action, next_timestep.reward,
                                 next_timestep.discount, extras))
    self._dangling = next_timestep.observation
    self._emit()
    if next_timestep.last():
      self._window.popleft()
      while self._window:
        self._emit()
        self._window.popleft()
      self.reset()

  def _emit(self):
    head = self._window[0]
    R, D = nstep_return([p.reward for p in self._window],
                        [p.discount for p in self._window], self._g)
    item = (head.observation, head.action, R, D, self._dangling)
    if head.extras:
      item = item + (head.extras,)
    w = self._writer()
    w.append(item)
    for table, fn in self._priority_fns.items():
      w.create_item(table=table, num_timesteps=1, priority=fn(None))
