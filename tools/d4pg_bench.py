"""D4PG learner updates/s on the control-suite-humanoid shape (67-d obs, 21-d act, C51 critic), batch 256, n=5
(BASELINE.json configs[2]); device-resident timing with CUDA events, whole step as a CUDA graph."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from acme_b200 import _capi, adders, d4pg, dm_env, loggers, networks, replay, specs


def main():
  prec = _capi.PRECISION_BF16 if (len(sys.argv) < 2 or sys.argv[1] != 'fp32') else _capi.PRECISION_FP32
  B, OBS, ACT, n, items = 256, 67, 21, 5, 100_000
  rng = np.random.default_rng(0)
  spec = specs.EnvironmentSpec(specs.Array((OBS,), np.float32), specs.BoundedArray((ACT,), np.float32, -1., 1.),
                               specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
  table = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Uniform(), replay.selectors.Fifo(), max_size=items,
                       rate_limiter=replay.rate_limiters.MinSize(1), signature=adders.NStepTransitionAdder.signature(spec),
                       max_window=n, discount=0.99, stage_slots=4096)
  server = replay.Server([table])
  adder = adders.NStepTransitionAdder(replay.Client(server), n_step=n, discount=0.99)
  steps = 0
  while steps < 20_000:     # 20k synthetic control steps through the product adder (setup, not timed)
    adder.add_first(dm_env.restart(rng.standard_normal(OBS).astype(np.float32)))
    T = 200
    for t in range(T):
      ts = dm_env.termination if t == T - 1 else dm_env.transition
      args = (np.float32(rng.random()), rng.standard_normal(OBS).astype(np.float32))
      adder.add(rng.uniform(-1, 1, ACT).astype(np.float32), ts(*args) if t == T - 1 else ts(args[0], args[1], np.float32(1.)))
      steps += 1
    table.flush()
  kw = dict(precision=prec) if 'precision' in networks.D4PGPolicy.__init__.__code__.co_varnames else {}
  policy, critic = networks.D4PGPolicy(OBS, ACT, seed=1, **kw), networks.D4PGCritic(OBS, ACT, seed=2, **kw)
  ds = replay.ReplayDataset(table, B, seed=11, stratified=False)
  L = d4pg.D4PGLearner(policy, critic, policy.clone(), critic.clone(), 0.99, target_update_period=100, dataset=ds,
                       logger=loggers.NoOpLogger())
  for _ in range(20):
    L.step(fetch_loss=False) if 'fetch_loss' in L.step.__code__.co_varnames else L.step()
  torch.cuda.synchronize()
  K = 300
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(K):
    L.step(fetch_loss=False) if 'fetch_loss' in L.step.__code__.co_varnames else L.step()
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / K
  out = {'workload': 'D4PG humanoid-shaped 67/21, C51 (51 atoms), batch 256, n=5', 'precision': 'fp32' if prec == 0 else 'tc',
         'updates_per_s': 1e3 / ms, 'ms_per_step': ms}
  print(json.dumps(out))
  os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
  json.dump(out, open(os.path.join(ROOT, 'gpurun_out', f"d4pg_bench_{out['precision']}.json"), 'w'))
  server.stop()


if __name__ == '__main__':
  main()
