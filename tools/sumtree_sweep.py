"""BASELINE config 4: prioritized sampling / update sweep over tree sizes and batch sizes on one GPU
(priorities only, no payload).  Reports samples/s, updates/s and the fraction of the measured HBM
peak using the ALGORITHMIC bytes of SURVEY §8d:
  sample: (L - S) * 128 + 28 bytes   (node lines below the staged levels + u, idx, key, prob)
  update: 4 + L * (128 + 4) + 12 bytes per updated item
Writes gpurun_out/sumtree_sweep.json (copy into profiles/ to keep)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from acme_b200 import replay

PEAK = 6523.3
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.exists(p):
  PEAK = json.load(open(p))['hbm_gbs']


def timeit(fn, iters):
  fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e-3


def main():
  sizes = [1 << 20, 1 << 22, 1 << 24, 100_000_000]
  batches = [256, 4096, 65536, 1 << 20]
  if len(sys.argv) > 1:
    sizes = [int(float(x)) for x in sys.argv[1].split(',')]
  out = []
  gen = torch.Generator(device='cuda')
  gen.manual_seed(0)
  for N in sizes:
    table = replay.Table.priorities_only('t', 0.6, N)
    w = torch.randn(N, device='cuda', generator=gen).abs_()
    table.set_weights(w)
    L, F, _ = table.tree_levels()
    widths = [len(table.read_tree_level(l)) for l in range(1, L)] + [N]
    for B in batches:
      u = torch.rand(B, device='cuda', generator=gen)
      idx = torch.empty(B, dtype=torch.int64, device='cuda')
      keys = torch.empty(B, dtype=torch.uint64, device='cuda')
      prob = torch.empty(B, device='cuda')
      pr = torch.randn(B, device='cuda', generator=gen).abs_()
      iters = 50 if B <= 65536 else 10
      ts = timeit(lambda: table.sample_into(u, idx, keys, prob, True), iters)
      tu = timeit(lambda: table.update_priorities_device(keys, pr), iters)
      # staged levels as chosen by the library: <= 16 KB for B < 16384, <= 160 KB otherwise
      budget = (160 if B >= 16384 else 16) * 1024
      used, S = 0, 0
      for l in range(1, L + 1):
        wl = widths[l - 1] * 4
        if used + wl > budget:
          break
        used += wl
        S = l
      sb = (L - S) * F * 4 + 28
      ub = 4 + L * (F * 4 + 4) + 12
      rec = dict(items=N, batch=B, levels=L, staged=S, sample_us=ts * 1e6, update_us=tu * 1e6,
                 samples_per_s=B / ts, updates_per_s=B / tu, sample_alg_bytes=sb, update_alg_bytes=ub,
                 sample_GBps=B * sb / ts / 1e9, update_GBps=B * ub / tu / 1e9,
                 sample_frac_hbm=B * sb / ts / 1e9 / PEAK, update_frac_hbm=B * ub / tu / 1e9 / PEAK)
      out.append(rec)
      print(f"N={N:>11,d} B={B:>8,d} L={L} S={S}  sample {ts*1e6:9.1f} us {B/ts/1e6:9.1f} Msamples/s {rec['sample_frac_hbm']*100:5.1f}% HBM |"
            f" update {tu*1e6:9.1f} us {B/tu/1e6:8.1f} Mupd/s {rec['update_frac_hbm']*100:5.1f}% HBM", flush=True)
    table.close()
    del w
  os.makedirs('gpurun_out', exist_ok=True)
  json.dump(dict(hbm_peak_gbs=PEAK, results=out), open('gpurun_out/sumtree_sweep.json', 'w'), indent=1)


main()
