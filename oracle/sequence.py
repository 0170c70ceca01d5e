"""Oracle: sequence items for recurrent learners (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, independently written:
  * `SequenceAdder`                       -> acme/adders/reverb/sequence.py:29-127
      (write every step; an item of `sequence_length` steps the first time the episode is that long and every `period`
      steps after; at the end of the episode a final step with zero action / reward / discount, then zero padding up to
      the next point where an item is due, then one more item check)
  * zero final step / zeros_like          -> acme/adders/reverb/utils.py:24-49
  * `compute_priority`                    -> acme/agents/tf/r2d2/learning.py:230-236
  * importance weights with an explicit N -> acme/agents/tf/r2d2/learning.py:170-176

Pinned by the reference's own 7 golden cases (`acme/adders/reverb/sequence_test.py:25-181` ->
`tests/golden/sequence_cases.json`, written by `tools/make_sequence_golden.py`).  The learner arithmetic
(compute_priority, importance weights) has no vectors in the reference: parity UNPINNED, restated from the cited lines.
"""

from __future__ import annotations

import collections
from typing import List, Tuple

import numpy as np

DEFAULT_TABLE = 'priority_table'  # acme/adders/reverb/base.py:30

SeqStep = collections.namedtuple('SeqStep', 'observation action reward discount start_of_episode extras')


def _zeros(x):
  if isinstance(x, (tuple, list)):
    return type(x)(_zeros(v) for v in x)
  if isinstance(x, dict):
    return {k: _zeros(v) for k, v in x.items()}
  return np.zeros_like(np.asarray(x))


def _is_due(count: int, length: int, period: int) -> bool:
  """sequence.py:111-117."""
  return count == length or (count > length and (count - length) % period == 0)


def end_padding(count: int, length: int, period: int) -> int:
  """sequence.py:95-102: zero steps appended after the final step (count includes the final step).  The reference feeds
  the value to range(), so a negative one (a long episode whose overshoot exceeds the period) pads nothing."""
  return max(0, length - count if count <= length else period - (count - length))


def enumerate_sequences(T: int, length: int, period: int, pad: bool = True) -> Tuple[int, List[int]]:
  """Closed form for one episode of T `add` calls: (steps written incl. final + padding, [first step of every item])."""
  starts = [k - length for k in range(1, T + 1) if _is_due(k, length, period)]
  count = T + 1
  if pad:
    count += end_padding(count, length, period)
  if _is_due(count, length, period):
    starts.append(count - length)
  return count, starts


class ReferenceSequenceAdder:
  """CPU restatement of `SequenceAdder` against a recording writer (oracle.nstep.RecordingClient)."""

  def __init__(self, client, sequence_length: int, period: int, pad_end_of_episode: bool = True, priority_fns=None):
    self._client = client
    self._length, self._period, self._pad = sequence_length, period, pad_end_of_episode
    self._window = collections.deque(maxlen=sequence_length)   # base.py:107 (buffer_size = sequence_length)
    self._dangling = None
    self._first = False
    self._count = 0
    self._writer_obj = None
    self._priority_fns = dict(priority_fns) if priority_fns else {DEFAULT_TABLE: lambda x: 1.}

  def _writer(self):
    if self._writer_obj is None:
      self._writer_obj = self._client.writer(self._length, delta_encoded=False, chunk_length=None)
    return self._writer_obj

  def reset(self):
    self._count = 0
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
    self._first = True

  def _put(self, step):
    self._window.append(step)
    self._writer().append(step)
    self._count += 1

  def _items_if_due(self):
    if not _is_due(self._count, self._length, self._period):
      return
    for table, fn in self._priority_fns.items():
      self._writer().create_item(table, len(self._window), fn(None))

  def add(self, action, next_timestep, extras=()):
    if self._dangling is None:
      raise ValueError('add_first must precede add')
    self._put(SeqStep(self._dangling, action, next_timestep.reward, next_timestep.discount, self._first, extras))
    self._dangling = next_timestep.observation
    self._first = False
    self._items_if_due()
    if next_timestep.last():
      head = self._window[0]
      final = SeqStep(self._dangling, _zeros(head.action), _zeros(head.reward), _zeros(head.discount), False,
                      _zeros(head.extras))
      self._put(final)
      if self._pad:
        blank = final._replace(observation=_zeros(final.observation))
        for _ in range(end_padding(self._count, self._length, self._period)):
          self._put(blank)
      self._items_if_due()
      self.reset()


# ----------------------------------------------------------------------------- learner arithmetic (r2d2/learning.py)
def compute_priority(errors, alpha):
  """:230-236 -- errors f32 [T, B]; tf.reduce_mean over axis 0 = fp32 sum in time order / T."""
  a = np.abs(np.asarray(errors, np.float32))
  mean = np.zeros(a.shape[1], np.float32)
  for t in range(a.shape[0]):
    mean = (mean + a[t]).astype(np.float32)
  mean = (mean / np.float32(a.shape[0])).astype(np.float32)
  mx = a.max(axis=0)
  al = np.float32(alpha)
  return (al * mx + (np.float32(1.) - al) * mean).astype(np.float32)


def importance_weights(probs, max_replay_size, exponent):
  """:170-176 -- f64: (1 / (N * p))^beta, divided by the max, cast to f32 where it multiplies the loss."""
  p = np.asarray(probs, np.float64)
  w = 1. / (float(max_replay_size) * p)
  w = w ** float(exponent)
  w = w / w.max()
  return w.astype(np.float32)
