"""Thread-safe hierarchical counter (semantics of `acme/utils/counting.py:27-119`)."""

import threading
import time
from typing import Dict, Optional


class Counter:

  def __init__(self, parent: Optional['Counter'] = None, prefix: str = '', time_delta: float = 1.0):
    self._parent, self._prefix, self._time_delta = parent, prefix, time_delta
    self._counts: Dict[str, float] = {}
    self._cache: Dict[str, float] = {}
    self._lock = threading.Lock()
    self._last_sync = 0.0

  def _prefixed(self, d):
    return {f'{self._prefix}_{k}': v for k, v in d.items()} if self._prefix else dict(d)

  def increment(self, **counts):
    with self._lock:
      for k, v in counts.items():
        self._counts[k] = self._counts.get(k, 0) + v
    return self.get_counts()

  def get_counts(self):
    now = time.time()
    if self._parent is not None and now - self._last_sync > self._time_delta:
      with self._lock:
        pending, self._counts = self._prefixed(self._counts), {}
      self._cache = self._parent.increment(**pending)
      self._last_sync = now
    out = self._prefixed(self._counts)
    for k, v in self._cache.items():
      out[k] = out.get(k, 0) + v
    return out

  def save(self):
    return {'counts': self._counts, 'cache': self._cache}

  def restore(self, state):
    self._last_sync = 0.
    self._counts, self._cache = state['counts'], state['cache']
