"""acme_b200 — the off-policy learner hot path of Acme (prioritized replay, n-step assembly,
DQN / D4PG learner step) rebuilt for NVIDIA B200 (sm_100a): hand-written CUDA behind a C ABI
(`include/b200rl.h`, `acme_b200/lib/libb200rl.so`), Python host code mirroring Acme's seams.

Importing the package works without a GPU (so host logic can be tested); any call that needs the
device raises if the library is missing or the GPU is not sm_100 — there is no CPU fallback.
"""

from acme_b200 import core, dm_env, specs  # noqa: F401
from acme_b200.agent import Agent  # noqa: F401
from acme_b200.core import Actor, Learner, Saveable, VariableSource, Worker  # noqa: F401
from acme_b200.environment_loop import EnvironmentLoop  # noqa: F401
from acme_b200.specs import EnvironmentSpec, make_environment_spec  # noqa: F401

__version__ = '0.1.0'
