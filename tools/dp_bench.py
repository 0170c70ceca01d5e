"""Times the fused reduce-scatter + Adam + all-gather kernel (parallel.PeerExchange.adam) under torchrun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from acme_b200 import parallel
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
dp = parallel.DataParallel(dist.group.WORLD)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_018_616
px = parallel.PeerExchange(dp, n, rank)
px.grads.normal_()
m, v = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
step = torch.zeros(1, dtype=torch.int64, device='cuda')
times = []
for it in range(12):
  dist.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  px.adam(0, n, m, v, step, 1e-3, 0.9, 0.999, 1e-8, 0, 0)
  e1.record(); torch.cuda.synchronize()
  step += 1
  times.append(e0.elapsed_time(e1) * 1e3)
px.check()
import ctypes, numpy as np
from acme_b200 import _capi
st = np.zeros(8, np.int64)
_capi.load().b200rl_debug_dp_stamps(px._h, ctypes.c_void_p(st.ctypes.data))
print(f'rank {rank}: wait-for-grads {(st[1]-st[0])/1e3:.1f} us, reduce+adam+broadcast {(st[2]-st[1])/1e3:.1f} us, final barrier {(st[3]-st[2])/1e3:.1f} us')
t = sorted(times[2:])[len(times[2:]) // 2]
shard = n / world * 4
if rank == 0:
  print(f'world {world} n {n}: fused exchange {t:.1f} us; peer read {shard * (world - 1) / 1e6:.1f} MB + peer write {shard * (world - 1) / 1e6:.1f} MB '
        f'-> {shard * (world - 1) / t / 1e3:.0f} GB/s each way; all {["%.0f" % x for x in times]}')
lib = _capi.load()
sink = torch.zeros(4, device='cuda')
peer = (rank + 1) % world
for mode, name in ((0, 'peer read'), (1, 'peer write'), (2, 'local read')):
  for blocks in (148, 296, 592, 1184):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.b200rl_debug_dp_probe(px._h, peer, ctypes.c_longlong(n), mode, blocks, ctypes.c_void_p(sink.data_ptr()), ctypes.c_void_p(_capi.current_stream()))
    e0.record()
    for _ in range(5):
      lib.b200rl_debug_dp_probe(px._h, peer, ctypes.c_longlong(n), mode, blocks, ctypes.c_void_p(sink.data_ptr()), ctypes.c_void_p(_capi.current_stream()))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 5
    if rank == 0: print(f'{name} {n * 4 / 1e6:.0f} MB, {blocks} CTAs: {us:.1f} us = {n * 4 / us / 1e3:.0f} GB/s')
dist.barrier()
px.close()
dist.destroy_process_group()
