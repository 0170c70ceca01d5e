"""Writes tests/golden/nstep_cases.json.

The seven cases are the golden vectors of the reference's own adder test,
`acme/adders/reverb/transition_test.py:29-170` (TEST_CASES), transcribed as data:
trajectory in -> list of expected (o, a, R, D, o'[, extras]) items out.  The reference
module cannot be imported in this image (dm_env / reverb / tensorflow absent), hence a
transcription rather than a dump.  Step kinds: "mid" = dm_env.transition (discount
defaults to 1.0), "term" = dm_env.termination (discount 0.0).
"""
import json
import os


def mid(r, o, d=1.0, extras=None):
  return dict(kind='mid', reward=r, observation=o, discount=d, extras=extras)


def term(r, o, extras=None):
  return dict(kind='term', reward=r, observation=o, discount=0.0, extras=extras)


CASES = [
    dict(name='OneStepFinalReward', n_step=1, additional_discount=1.0, first=1,
         steps=[mid(0.0, 2), mid(0.0, 3), term(1.0, 4)],
         expected=[[1, 0, 0.0, 1.0, 2], [2, 0, 0.0, 1.0, 3], [3, 0, 1.0, 0.0, 4]]),
    dict(name='OneStepDict', n_step=1, additional_discount=1.0, first={'foo': 1},
         steps=[mid(0.0, {'foo': 2}), mid(0.0, {'foo': 3}), term(1.0, {'foo': 4})],
         expected=[[{'foo': 1}, 0, 0.0, 1.0, {'foo': 2}], [{'foo': 2}, 0, 0.0, 1.0, {'foo': 3}],
                   [{'foo': 3}, 0, 1.0, 0.0, {'foo': 4}]]),
    dict(name='OneStepExtras', n_step=1, additional_discount=1.0, first=1,
         steps=[mid(0.0, 2, extras={'state': 0}), mid(0.0, 3, extras={'state': 1}),
                term(1.0, 4, extras={'state': 2})],
         expected=[[1, 0, 0.0, 1.0, 2, {'state': 0}], [2, 0, 0.0, 1.0, 3, {'state': 1}],
                   [3, 0, 1.0, 0.0, 4, {'state': 2}]]),
    dict(name='TwoStep', n_step=2, additional_discount=1.0, first=1,
         steps=[mid(1.0, 2, 0.5), mid(1.0, 3, 0.5), term(1.0, 4)],
         expected=[[1, 0, 1.0, 0.50, 2], [1, 0, 1.5, 0.25, 3], [2, 0, 1.5, 0.00, 4], [3, 0, 1.0, 0.00, 4]]),
    dict(name='TwoStepWithExtras', n_step=2, additional_discount=1.0, first=1,
         steps=[mid(1.0, 2, 0.5, extras={'state': 0}), mid(1.0, 3, 0.5, extras={'state': 1}),
                term(1.0, 4, extras={'state': 2})],
         expected=[[1, 0, 1.0, 0.50, 2, {'state': 0}], [1, 0, 1.5, 0.25, 3, {'state': 0}],
                   [2, 0, 1.5, 0.00, 4, {'state': 1}], [3, 0, 1.0, 0.00, 4, {'state': 2}]]),
    dict(name='ThreeStepDiscounted', n_step=3, additional_discount=0.4, first=1,
         steps=[mid(1.0, 2, 0.5), mid(1.0, 3, 0.5), term(1.0, 4)],
         expected=[[1, 0, 1.00, 0.5, 2], [1, 0, 1.20, 0.1, 3], [1, 0, 1.24, 0.0, 4],
                   [2, 0, 1.20, 0.0, 4], [3, 0, 1.00, 0.0, 4]]),
    dict(name='ThreeStepVaryingReward', n_step=3, additional_discount=0.5, first=1,
         steps=[mid(2.0, 2), mid(3.0, 3), mid(5.0, 4), term(7.0, 5)],
         expected=[[1, 0, 2, 1.00, 2], [1, 0, 2 + 0.5 * 3, 0.50, 3], [1, 0, 2 + 0.5 * 3 + 0.25 * 5, 0.25, 4],
                   [2, 0, 3 + 0.5 * 5 + 0.25 * 7, 0.00, 5], [3, 0, 5 + 0.5 * 7, 0.00, 5], [4, 0, 7, 0.00, 5]]),
]

if __name__ == '__main__':
  path = os.path.join(os.path.dirname(__file__), 'nstep_cases.json')
  with open(path, 'w') as f:
    json.dump(CASES, f, indent=1)
  print('wrote', path, len(CASES), 'cases')
