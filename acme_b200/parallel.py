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
