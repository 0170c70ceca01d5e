"""world_size-2 `gloo` tests (CPU) of the data-parallel contract used by the sharded learner
(ordering of acme/agents/tf/crr/recurrent_learning.py:346-359: all-reduce-mean the
gradients, then clip / apply): per-rank half batches + global importance-weight max + summed
gradients x 1/R must reproduce the single-process gradient of the full batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _make_batch(seed, B, obs_dim, A):
  rng = np.random.default_rng(seed)
  return dict(o0=rng.standard_normal((B, obs_dim)).astype(np.float32), o1=rng.standard_normal((B, obs_dim)).astype(np.float32),
              a=rng.integers(0, A, B), R=rng.standard_normal(B).astype(np.float32),
              D=rng.choice([0., 1.], B).astype(np.float32), prob=rng.uniform(1e-6, 1e-3, B).astype(np.float32))


def _grads(net, tgt, batch, wmax, scale_B):
  """d(mean loss)/d(params) of one (sub)batch with a given importance-weight normaliser."""
  from oracle import losses
  o0, o1 = torch.tensor(batch['o0']), torch.tensor(batch['o1'])
  q_tm1 = net(o0)
  with torch.no_grad():
    q_tv, q_ts = tgt(o1), net(o1)
  ref = losses.dqn_loss(q_tm1.detach().numpy(), q_tv.numpy(), q_ts.numpy(), batch['a'], batch['R'], batch['D'],
                        batch['prob'], 0.99, 1.0, 0.2, 1.0, global_wmax=wmax)
  names = net.names()
  g = torch.autograd.grad(q_tm1, [net.vars[k] for k in names], grad_outputs=torch.tensor(ref['dq_tm1']))
  return {k: x.numpy() for k, x in zip(names, g)}, ref


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from acme_b200 import parallel
    from oracle import nets as onets
    dp = parallel.DataParallel(dist.group.WORLD)
    assert dp.world == world and dp.rank == rank and dp.grad_scale == 1.0 / world
    B, obs_dim, A = 16, 12, 5
    net, tgt = onets.MLPQNetwork(obs_dim, [32, A], seed=0), onets.MLPQNetwork(obs_dim, [32, A], seed=1)
    full = _make_batch(7, B * world, obs_dim, A)
    mine = {k: v[rank * B:(rank + 1) * B] for k, v in full.items()}
    # 1) global normaliser = all-reduce(MAX) of the local maxima (f64)
    local_max = torch.tensor([((1.0 / mine['prob'].astype(np.float64))**0.2).max()], dtype=torch.float64)
    wmax = float(dp.global_max_(local_max.clone())[0])
    assert wmax == ((1.0 / full['prob'].astype(np.float64))**0.2).max()
    # 2) local gradient of the local mean, summed over ranks, x 1/R
    g, _ = _grads(net, tgt, mine, wmax, B)
    flat = torch.cat([torch.tensor(g[k]).reshape(-1) for k in net.names()])
    dp.sum_(flat)
    flat *= dp.grad_scale
    # reference: one process, the whole batch, mean over world * B samples
    g_full, _ = _grads(net, tgt, full, wmax, B * world)
    want = np.concatenate([g_full[k].reshape(-1) for k in net.names()])
    np.testing.assert_allclose(flat.numpy(), want, rtol=1e-5, atol=1e-7)
    # 3) shard masses and the replication check
    masses = dp.gather_scalars(torch.tensor(float(rank + 1)))
    assert masses.tolist() == [1.0, 2.0]
    dp.assert_replicated(torch.ones(4))
    try:
      dp.assert_replicated(torch.full((4,), float(rank)))
      diverged = False
    except RuntimeError:
      diverged = True
    assert diverged
    out[rank] = 'ok'
  finally:
    dist.destroy_process_group()


def test_two_rank_data_parallel_gradient_matches_single_process():
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  assert dict(out) == {0: 'ok', 1: 'ok'}


def test_sharded_probability_is_unbiased():
  """Rank r reports q_i = w_i / (R * M_r): with every rank drawing B/R items the expected
  number of draws of item i is exactly B * q_i, so importance weights 1/q_i stay unbiased."""
  from oracle import replay as oreplay
  rng = np.random.default_rng(0)
  R, n = 2, 64
  for r in range(R):
    t = oreplay.Table(n, 2 * n + 8, (1,), np.float32, (), np.int32, 0.99, 1.0, max_window=1, shard_count=R)
    w = t.writer()
    for i in range(n):
      t.append(w, [0.], 0, 0., 1., [0.])
      t.create_item(w, 1, float(rng.uniform(0.1, 2.0)))
    u = rng.random(4096, dtype=np.float32)
    keys, pos, prob = t.sample(u, stratified=True)
    leaves = t.tree.leaves[:n]
    np.testing.assert_allclose(prob, leaves[pos] / (R * t.tree.total), rtol=1e-6)
    freq = np.bincount(pos, minlength=n) / len(u)
    np.testing.assert_allclose(freq, leaves / t.tree.total, atol=5e-3)
