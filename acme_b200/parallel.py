"""Data-parallel plumbing for the learner (one process per GPU, torch.distributed).

The reference's DQN / D4PG learners contain no collective; its only multi-replica learner (CRR) does
`all_reduce('mean', grads)`, then clips, then applies (`acme/agents/tf/crr/recurrent_learning.py:346-359`).
This module keeps that ordering for the sharded-replay data-parallel learner:
  * every rank owns a replay shard and samples B/R items from it;
  * one scalar all-reduce(MAX) makes the importance-weight normaliser global (the reference divides by the
    batch max, `dqn/learning.py:140`);
  * the flat fp32 gradient buffer is all-reduced (SUM) and Adam multiplies by 1/R: parameters stay replicated.
Works on any backend: `nccl` on GPUs (NVLink / NVSwitch), `gloo` in the CPU tests.
"""

from __future__ import annotations

import os
from typing import Optional


class DataParallel:

  def __init__(self, process_group=None):
    self.group = process_group
    self.world, self.rank = 1, 0
    if process_group is not None:
      import torch.distributed as dist
      self.world = dist.get_world_size(process_group)
      self.rank = dist.get_rank(process_group)

  @property
  def enabled(self) -> bool:
    return self.world > 1

  @property
  def grad_scale(self) -> float:
    """all-reduce(SUM) then x 1/R == all-reduce('mean')."""
    return 1.0 / self.world

  def global_max_(self, scalar_tensor):
    if self.enabled:
      import torch.distributed as dist
      dist.all_reduce(scalar_tensor, op=dist.ReduceOp.MAX, group=self.group)
    return scalar_tensor

  def sum_(self, tensor):
    if self.enabled:
      import torch.distributed as dist
      dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
    return tensor

  def gather_scalars(self, scalar_tensor):
    """all-gather of one scalar per rank (shard masses M_r for monitoring shard imbalance)."""
    import torch
    if not self.enabled:
      return scalar_tensor.reshape(1).clone()
    import torch.distributed as dist
    out = torch.empty(self.world, dtype=scalar_tensor.dtype, device=scalar_tensor.device)
    dist.all_gather_into_tensor(out, scalar_tensor.reshape(1), group=self.group)
    return out

  def refresh_global_mass(self, table, total=None):
    """Global-priority-mass normalisation: all-reduce(SUM) of the shards' masses into `total` (device float
    [1], allocated on first use) and install it in the table, whose K1 then reports weight / sum_r M_r.  Call after
    priority updates / inserts, whenever the reported probabilities should track the global mass."""
    import torch
    if total is None:
      total = torch.empty(1, dtype=torch.float32, device=torch.device('cuda', table.device))
    total.copy_(table.mass_tensor())
    if self.enabled:
      import torch.distributed as dist
      dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
    table.set_global_mass(total)
    return total

  def assert_replicated(self, tensor, what: str = 'parameters'):
    """Raises if ranks have diverged (cheap: two scalar all-reduces of a checksum)."""
    if not self.enabled:
      return
    import torch
    import torch.distributed as dist
    s = tensor.double().sum().reshape(1)
    hi, lo = s.clone(), s.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
    if not torch.equal(hi, lo):
      raise RuntimeError(f'{what} differ across data-parallel ranks: checksum range [{lo.item()}, {hi.item()}]')


class _RawDeviceArray:
  """A device pointer dressed up for torch.as_tensor (zero copy)."""

  def __init__(self, ptr: int, n: int):
    self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f4', 'data': (ptr, False), 'version': 2}


class PeerExchange:
  """The learner's exchange over NVLink peer memory (`include/b200rl.h` "data-parallel learner"): one fused
  reduce-scatter + Adam + all-gather kernel per gradient bucket, and the scalar all-reduce(MAX).

  The flat parameter and gradient buffers live in a region that every peer maps through CUDA IPC; `torch.distributed`
  is used once, to exchange the 64-byte handles.  Requires one process per GPU on one node with P2P access.
  """

  def __init__(self, dp: DataParallel, n_params: int, device: int):
    import ctypes
    import torch
    import torch.distributed as dist
    from acme_b200 import _capi
    assert dp.enabled
    self.dp, self.n = dp, int(n_params)
    self._h = ctypes.c_void_p()
    dev = torch.device('cuda', device)
    cfg = _capi.DpCfg(world=dp.world, rank=dp.rank, device=device, reserved=0, n_params=self.n)
    self._symm = None
    self.multicast = False
    if os.environ.get('B200RL_DP_SYMM', '1') != '0':
      # Symmetric allocation with an NVSwitch multicast mapping (torch symmetric memory supplies the allocation and the
      # rendezvous, nothing on the data path): enables the in-switch reduction of `adam_mc`.
      try:
        import torch.distributed._symmetric_memory as symm
        nbytes = int(_capi.load().b200rl_dp_region_bytes(self.n))
        buf = symm.empty(nbytes, dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(buf, dp.group.group_name if dp.group is not None else dist.group.WORLD.group_name)
        bases = (ctypes.c_void_p * dp.world)(*[int(p) for p in hdl.buffer_ptrs])
        mc = int(hdl.multicast_ptr) if getattr(hdl, 'multicast_ptr', 0) else None
        _capi.call('b200rl_dp_create_external', ctypes.byref(self._h), ctypes.byref(cfg), bases, mc)
        self._symm = (buf, hdl)
        self.multicast = bool(_capi.load().b200rl_dp_has_multicast(self._h))
      except Exception as e:  # noqa: BLE001 -- no symmetric memory on this box / torch build: CUDA-IPC region instead
        if os.environ.get('B200RL_DP_SYMM') == '1':
          raise
        self._symm, self._h = None, ctypes.c_void_p()
        self._symm_error = repr(e)
    if self._symm is None:
      _capi.call('b200rl_dp_create', ctypes.byref(self._h), ctypes.byref(cfg))
    p, g = ctypes.c_void_p(), ctypes.c_void_p()
    _capi.call('b200rl_dp_buffers', self._h, ctypes.byref(p), ctypes.byref(g))
    self.params = torch.as_tensor(_RawDeviceArray(p.value, self.n), device=dev)
    self.grads = torch.as_tensor(_RawDeviceArray(g.value, self.n), device=dev)
    self._max_epoch = torch.zeros(1, dtype=torch.int64, device=dev)
    if self._symm is None:
      mine = ctypes.create_string_buffer(64)
      _capi.call('b200rl_dp_export', self._h, mine)
      handles = [None] * dp.world
      dist.all_gather_object(handles, bytes(mine.raw), group=dp.group)
      for r, hb in enumerate(handles):
        if r != dp.rank:
          _capi.call('b200rl_dp_import', self._h, r, ctypes.create_string_buffer(hb, 64))
    torch.cuda.synchronize()
    dist.barrier(group=dp.group)       # every region is zeroed and mapped before anyone's first exchange

  def max_f64_(self, value, step=None):
    """all-reduce(MAX) of one f64 through the peers' mailboxes.  The barrier epoch is a device counter owned by this
    object and advanced by every call (graph-capturable), so consecutive exchanges never share an epoch whatever the
    caller's own counters do (injected uniforms do not advance the dataset's draw counter; restore() can rewind the
    step counter).  `step` is accepted for backward compatibility and ignored."""
    from acme_b200 import _capi
    st = _capi.current_stream()
    _capi.call('b200rl_dp_max_f64', self._h, _capi.ptr(value), _capi.ptr(self._max_epoch), st)
    _capi.call('b200rl_step_increment', _capi.ptr(self._max_epoch), st)

  def adam(self, off: int, n: int, m, v, step, lr: float, b1: float, b2: float, eps: float, eps_mode: int, bucket: int,
           final_barrier: bool = True):
    from acme_b200 import _capi
    _capi.call('b200rl_dp_adam', self._h, off, n, _capi.ptr(m), _capi.ptr(v), _capi.ptr(step), lr, b1, b2, eps, eps_mode,
               bucket, int(final_barrier), _capi.current_stream())

  def adam_mc(self, off: int, n: int, m, v, step, lr: float, b1: float, b2: float, eps: float, eps_mode: int, bucket: int,
              final_barrier: bool = True, max_ctas: int = 0):
    """`adam` with the reduction done inside the NVSwitch (`b200rl_dp_adam_mc`); needs `self.multicast`."""
    from acme_b200 import _capi
    _capi.call('b200rl_dp_adam_mc', self._h, off, n, _capi.ptr(m), _capi.ptr(v), _capi.ptr(step), lr, b1, b2, eps, eps_mode,
               bucket, int(final_barrier), max_ctas, _capi.current_stream())

  def reduce_adam_mc(self, off: int, n: int, m, v, step, lr: float, b1: float, b2: float, eps: float, eps_mode: int,
                     bucket: int, shadow_ptr=None, max_ctas: int = 0):
    from acme_b200 import _capi
    _capi.call('b200rl_dp_reduce_adam_mc', self._h, off, n, _capi.ptr(m), _capi.ptr(v), _capi.ptr(step), lr, b1, b2, eps,
               eps_mode, bucket, shadow_ptr, max_ctas, _capi.current_stream())

  def broadcast_mc(self, off: int, n: int, step, bucket: int, final_barrier: bool = True, max_ctas: int = 0):
    from acme_b200 import _capi
    _capi.call('b200rl_dp_broadcast_mc', self._h, off, n, _capi.ptr(step), bucket, int(final_barrier), max_ctas,
               _capi.current_stream())

  def reduce_adam_ce(self, off: int, n: int, m, v, step, lr: float, b1: float, b2: float, eps: float, eps_mode: int,
                     bucket: int, shadow_ptr=None, max_ctas: int = 0):
    """First half of the copy-engine exchange (`b200rl_dp_reduce_adam_ce`): wait for the peers' gradients, pull the owned
    shard by DMA, Adam on it (parameters + bf16 shadow of this rank only)."""
    from acme_b200 import _capi
    _capi.call('b200rl_dp_reduce_adam_ce', self._h, off, n, _capi.ptr(m), _capi.ptr(v), _capi.ptr(step), lr, b1, b2, eps,
               eps_mode, bucket, shadow_ptr, max_ctas, _capi.current_stream())

  def broadcast_ce(self, off: int, n: int, step, bucket: int, final_barrier: bool = True):
    """Second half: push the owned shard of the new parameters to every peer by DMA, then the barrier."""
    from acme_b200 import _capi
    _capi.call('b200rl_dp_broadcast_ce', self._h, off, n, _capi.ptr(step), bucket, int(final_barrier),
               _capi.current_stream())

  def check(self):
    from acme_b200 import _capi
    _capi.call('b200rl_dp_status', self._h)

  def close(self):
    import torch
    import torch.distributed as dist
    from acme_b200 import _capi
    if self._h:
      torch.cuda.synchronize()
      dist.barrier(group=self.dp.group)     # nobody unmaps while a peer may still touch the region
      self.params = self.grads = None
      _capi.call('b200rl_dp_destroy', self._h)
      self._h = None
      self._symm = None
