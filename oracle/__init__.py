"""CPU oracle for the off-policy learner hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain NumPy / PyTorch-CPU, the arithmetic of the
reference's hot path (tmtlakmal/acme v0.1.8) so that
the CUDA product in `acme_b200/` has something to be checked against.

Rules (enforced by `tests/test_boundary.py::test_product_does_not_import_oracle`):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
    `--impl reference` legs may import anything from here;
  * nothing in `acme_b200/` imports it; the product has no CPU fallback.

Parity status per module:
  * `oracle.nstep`   — PINNED: reproduces the 7 golden cases of
    `acme/adders/reverb/transition_test.py:29-170` (tests/test_oracle_nstep.py,
    fixtures in tests/golden/nstep_cases.json).
  * `oracle.sumtree`, `oracle.replay` — parity UNPINNED: Reverb
    (dm-reverb-nightly==0.1.0.dev20200708, `setup.py:30`) is not in the reference
    tree; the module restates its published Prioritized/Fifo semantics as used at
    `acme/agents/tf/dqn/agent.py:95-101`, checked by closed-form properties.
  * `oracle.losses`  — Huber / l2_project / dpg are transcriptions of
    `acme/tf/losses/{huber,distributional,dpg}.py` (files present), but the
    reference holds no tests for them and trfl / TF are not installable:
    parity UNPINNED beyond the cited lines, checked against closed forms and an
    independent PyTorch autograd evaluation.
  * `oracle.nets`, `oracle.learner` — Sonnet / TF layer semantics restated from
    memory of their public sources (SURVEY.md §8c caveat): parity UNPINNED.
"""
