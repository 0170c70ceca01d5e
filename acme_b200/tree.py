"""Minimal nest helpers (the subset of dm-tree the hot-path seams use: `adders/reverb/base.py:28`)."""

from typing import Any, Callable, List


def _is_namedtuple(x) -> bool:
  return isinstance(x, tuple) and hasattr(x, '_fields')


def is_nest(x) -> bool:
  return isinstance(x, (dict, list, tuple))


def flatten(x) -> List[Any]:
  """Leaves in deterministic order (dict keys sorted, like dm-tree)."""
  if isinstance(x, dict):
    out = []
    for k in sorted(x):
      out.extend(flatten(x[k]))
    return out
  if isinstance(x, (list, tuple)):
    out = []
    for v in x:
      out.extend(flatten(v))
    return out
  return [x]


def unflatten_as(structure, leaves: List[Any]):
  it = iter(leaves)

  def build(s):
    if isinstance(s, dict):
      return type(s)((k, build(s[k])) for k in sorted(s))
    if _is_namedtuple(s):
      return type(s)(*[build(v) for v in s])
    if isinstance(s, (list, tuple)):
      return type(s)(build(v) for v in s)
    return next(it)

  return build(structure)


def map_structure(fn: Callable, *structures):
  flat = [flatten(s) for s in structures]
  n = len(flat[0])
  if any(len(f) != n for f in flat):
    raise ValueError('structures do not match')
  return unflatten_as(structures[0], [fn(*xs) for xs in zip(*flat)])
