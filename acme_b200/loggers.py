"""`Logger.write(dict)` seam (`acme/utils/loggers/base.py:27-32`) with the loggers the hot path
constructs by default (terminal with a time filter, `loggers/terminal.py:62-92`, `filters.py`)."""

import abc
import time
from typing import Any, Callable, Mapping


class Logger(abc.ABC):

  @abc.abstractmethod
  def write(self, data: Mapping[str, Any]):
    ...


class NoOpLogger(Logger):

  def write(self, data):
    pass


class InMemoryLogger(Logger):

  def __init__(self):
    self.data = []

  def write(self, data):
    self.data.append(dict(data))


def _fmt(v):
  try:
    return f'{float(v):0.3f}'
  except Exception:  # noqa: BLE001
    return str(v)


class TerminalLogger(Logger):

  def __init__(self, label: str = '', print_fn: Callable[[str], None] = print, time_delta: float = 0.0):
    self._label, self._print, self._dt, self._t = label, print_fn, time_delta, 0.0

  def write(self, data):
    now = time.time()
    if now - self._t < self._dt:
      return
    self._t = now
    body = ' | '.join(f'{k.replace("_", " ").title()} = {_fmt(v)}' for k, v in sorted(data.items()))
    self._print(f'[{self._label.title()}] {body}' if self._label else body)


def make_default_logger(label: str, time_delta: float = 1.0) -> Logger:
  return TerminalLogger(label, time_delta=time_delta)
