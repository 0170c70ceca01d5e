"""Does torch's symmetric memory give a multicast (NVLS) mapping on this box?  torchrun --nproc-per-node N tools/symm_probe.py"""
import os, torch, torch.distributed as dist
rank = int(os.environ['RANK']); torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
import torch.distributed._symmetric_memory as symm
try:
  t = symm.empty(1 << 22, dtype=torch.float32, device=torch.device('cuda', rank))
  h = symm.rendezvous(t, dist.group.WORLD.group_name)
  print(rank, 'buffer_ptrs', [hex(p) for p in h.buffer_ptrs], 'multicast_ptr', hex(h.multicast_ptr), 'signal_pads', [hex(p) for p in h.signal_pad_ptrs],
        'buffer_size', h.buffer_size, 'signal_pad_size', h.signal_pad_size, flush=True)
  t.fill_(rank + 1.0)
  dist.barrier(); torch.cuda.synchronize()
  if h.multicast_ptr:
    out = torch.ops.symm_mem.multimem_all_reduce_(t, 'sum', dist.group.WORLD.group_name)
    torch.cuda.synchronize()
    print(rank, 'multimem all-reduce ->', float(t[0]), float(t[-1]), flush=True)
except Exception as e:
  print(rank, 'symmetric memory failed:', repr(e), flush=True)
dist.barrier()
dist.destroy_process_group()
