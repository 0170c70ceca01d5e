"""The HBM-bound kernels of the hot path, one launch group each, for `ncu --set full`:
  K1 sample    N = 1e8 items, B = 2^20 (HBM regime) and B = 256 (the learner's call)
  K2 update    same trees, same batches
  K3 gather    uint8 84x84x4 ring, B = 256, n = 3
  K5 C51       projection + CE, K = 51, B = 256 and B = 65,536
  K7 Adam      8,018,611 parameters
Each group is also timed with CUDA events (warm, 20 iterations) and written to gpurun_out/hbm_kernels.json with the
algorithmic bytes of SURVEY §8d.  Usage under ncu:
  python tools/ncu_hbm_kernels.py --once && ncu --set full --clock-control none --import-source on \
      --profile-from-start off -o gpurun_out/hbm_r02 python tools/ncu_hbm_kernels.py --once
(--once: each group is launched once, warm, inside a cudaProfilerStart/Stop window)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from acme_b200 import _capi, adders, replay, specs

ONCE = '--once' in sys.argv
PEAK = bench.peaks()['hbm']


def timeit(fn, iters=20):
  fn()
  torch.cuda.synchronize()
  if ONCE:            # one warm launch group inside the profiler window (ncu --profile-from-start off)
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    return float('nan')
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e-3


def main():
  gen = torch.Generator(device='cuda')
  gen.manual_seed(0)
  out = {}
  # ---- K1 / K2 on a priorities-only tree
  N = 100_000_000
  table = replay.Table.priorities_only('t', 0.6, N)
  w = torch.randn(N, device='cuda', generator=gen).abs_()
  table.set_weights(w)
  L, F, S = table.tree_levels()
  for B in (1 << 20, 256):
    u = torch.rand(B, device='cuda', generator=gen)
    idx = torch.empty(B, dtype=torch.int64, device='cuda')
    keys = torch.empty(B, dtype=torch.uint64, device='cuda')
    prob = torch.empty(B, device='cuda')
    pr = torch.randn(B, device='cuda', generator=gen).abs_()
    ts = timeit(lambda: table.sample_into(u, idx, keys, prob, True))
    tu = timeit(lambda: table.update_priorities_device(keys, pr))
    out[f'k1_sample_N1e8_B{B}'] = dict(us=ts * 1e6, levels=L, fanout=F)
    out[f'k2_update_N1e8_B{B}'] = dict(us=tu * 1e6)
  table.close()
  del w
  torch.cuda.empty_cache()
  # ---- K3 on a payload ring (150k steps = 4.2 GB: far larger than L2)
  items = 150_000
  spec = specs.EnvironmentSpec(specs.Array(bench.OBS_SHAPE, np.uint8), specs.DiscreteArray(bench.NUM_ACTIONS),
                               specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
  t2 = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Prioritized(0.6), replay.selectors.Fifo(), max_size=items,
                    rate_limiter=replay.rate_limiters.MinSize(1), signature=adders.NStepTransitionAdder.signature(spec),
                    max_window=3, discount=0.99, slot_capacity=items + 4096, stage_slots=4096)
  bench.fill_replay(t2, items, 3, seed=3)
  ds = replay.ReplayDataset(t2, 256, seed=1)
  ds.sample_only()
  tg = timeit(ds.gather_only)
  out['k3_gather_B256'] = dict(us=tg * 1e6, alg_bytes=256 * 112_944, GBps=256 * 112_944 / tg / 1e9 if tg == tg else None)
  t2.close()
  # ---- K5
  for B in (256, 65536):
    K = 51
    lt1, lt = torch.randn(B, K, device='cuda', generator=gen), torch.randn(B, K, device='cuda', generator=gen)
    R, D = torch.rand(B, device='cuda', generator=gen), torch.ones(B, device='cuda')
    tgt, lps, dl = torch.empty(B, K, device='cuda'), torch.empty(B, device='cuda'), torch.empty(B, K, device='cuda')
    t5 = timeit(lambda: _capi.call('b200rl_c51_loss', B, K, -150., 150., lt1.data_ptr(), lt.data_ptr(), R.data_ptr(), D.data_ptr(),
                                   0.99, 1.0 / B, tgt.data_ptr(), lps.data_ptr(), dl.data_ptr(), None, _capi.current_stream()))
    out[f'k5_c51_B{B}'] = dict(us=t5 * 1e6, alg_bytes=B * 828, GBps=B * 828 / t5 / 1e9 if t5 == t5 else None)
  # ---- K7
  n = 8_018_624
  p, g, m, v = (torch.randn(n, device='cuda', generator=gen) for _ in range(4))
  v.abs_()
  step = torch.zeros(1, dtype=torch.int64, device='cuda')
  t7 = timeit(lambda: _capi.call('b200rl_adam', n, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), step.data_ptr(), 1e-3,
                                 0.9, 0.999, 1e-8, 0, None, None, _capi.current_stream()))
  out['k7_adam_8M'] = dict(us=t7 * 1e6, alg_bytes=n * 28, GBps=n * 28 / t7 / 1e9 if t7 == t7 else None)
  for k, r in out.items():
    if r.get('GBps'):
      r['frac_of_measured_hbm_peak'] = r['GBps'] / PEAK
  if not ONCE:
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(dict(hbm_peak_gbs=PEAK, results=out), open(os.path.join(ROOT, 'gpurun_out', 'hbm_kernels.json'), 'w'), indent=1)
    print(json.dumps(out, indent=1))


main()
