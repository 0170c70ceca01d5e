"""Environment specs: the four-field bundle agents are built from (`acme/specs.py:26-49`).

The field names and the `make_environment_spec(environment)` helper are the reference's API; the spec classes themselves
come from this package's `dm_env` shim (the real `dm_env` is not installable offline).
"""

import collections

from acme_b200.dm_env import specs as _dm_specs

Array, BoundedArray, DiscreteArray = _dm_specs.Array, _dm_specs.BoundedArray, _dm_specs.DiscreteArray

# (field of the bundle, environment method that provides it)
_SOURCES = (('observations', 'observation_spec'), ('actions', 'action_spec'),
            ('rewards', 'reward_spec'), ('discounts', 'discount_spec'))

EnvironmentSpec = collections.namedtuple('EnvironmentSpec', [field for field, _ in _SOURCES])


def make_environment_spec(environment) -> EnvironmentSpec:
  return EnvironmentSpec(**{field: getattr(environment, method)() for field, method in _SOURCES})
