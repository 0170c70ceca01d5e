"""Data-parallel learner over NVLink peer memory (`acme_b200/csrc/dp_p2p.cu`): two ranks, two GPUs.

Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2`).  Checks the fused reduce-scatter + Adam +
all-gather against NCCL all-reduce + the replicated Adam kernel on identical gradients, the scalar all-reduce(MAX),
and a few whole learner steps (replicas must stay bit-identical; losses must match the NCCL path).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import torch
  import torch.distributed as dist
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
  try:
    from acme_b200 import _capi, parallel
    dp = parallel.DataParallel(dist.group.WORLD)
    n = 100_003 * 4
    px = parallel.PeerExchange(dp, n, rank)
    px_ce = parallel.PeerExchange(dp, n, rank)       # the copy-engine form of the same exchange, fed the same gradients
    px_mc = parallel.PeerExchange(dp, n, rank)       # the multicast (in-switch reduction) form, where the box has NVLS
    if not px_mc.multicast:
      px_mc.close()
      px_mc = None
    gen = torch.Generator(device='cuda').manual_seed(7)            # same parameters everywhere
    p0 = torch.randn(n, device='cuda', generator=gen)
    gen_r = torch.Generator(device='cuda').manual_seed(100 + rank)  # rank-specific gradients
    step = torch.zeros(1, dtype=torch.int64, device='cuda')
    px.params.copy_(p0)
    px_ce.params.copy_(p0)
    if px_mc is not None:
      px_mc.params.copy_(p0)
      m_mc, v_mc = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    m, v = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    m_ce, v_ce = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    shadow_ce = torch.zeros(n, dtype=torch.bfloat16, device='cuda')
    p_ref, m_ref, v_ref = p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    gscale = torch.full((1,), 1.0 / world, device='cuda')
    for it in range(3):
      g = torch.randn(n, device='cuda', generator=gen_r)
      px.grads.copy_(g)
      # reference: NCCL all-reduce(SUM) then the replicated Adam kernel with 1/R
      g_sum = g.clone()
      dist.all_reduce(g_sum)
      _capi.call('b200rl_adam', n, _capi.ptr(p_ref), _capi.ptr(g_sum), _capi.ptr(m_ref), _capi.ptr(v_ref), _capi.ptr(step),
                 1e-3, 0.9, 0.999, 1e-8, 0, _capi.ptr(gscale), None, _capi.current_stream())
      # two buckets, odd split, like the learner
      cut = 40_000
      px.adam(cut, n - cut, m, v, step, 1e-3, 0.9, 0.999, 1e-8, 0, 0)
      px.adam(0, cut, m, v, step, 1e-3, 0.9, 0.999, 1e-8, 0, 1)
      # copy-engine form: reduce + Adam half, then the broadcast half; bit-identical to the SM-issued kernel
      px_ce.grads.copy_(g)
      for (o, k, bucket) in ((cut, n - cut, 0), (0, cut, 1)):
        px_ce.reduce_adam_ce(o, k, m_ce, v_ce, step, 1e-3, 0.9, 0.999, 1e-8, 0, bucket, shadow_ce.data_ptr(), 64 if bucket == 0 else 0)
      for (o, k, bucket) in ((cut, n - cut, 0), (0, cut, 1)):
        px_ce.broadcast_ce(o, k, step, bucket, final_barrier=True)
      torch.cuda.synchronize()
      px_ce.check()
      assert torch.equal(px_ce.params, px.params), (it, (px_ce.params - px.params).abs().max().item())
      # in-switch reduction (multimem): the sum order differs from rank order by an ulp of the gradient sum
      if px_mc is not None:
        px_mc.grads.copy_(g)
        torch.cuda.synchronize(); dist.barrier()
        px_mc.adam_mc(cut, n - cut, m_mc, v_mc, step, 1e-3, 0.9, 0.999, 1e-8, 0, 0, final_barrier=False, max_ctas=16)
        px_mc.adam_mc(0, cut, m_mc, v_mc, step, 1e-3, 0.9, 0.999, 1e-8, 0, 1, final_barrier=True)
        torch.cuda.synchronize(); dist.barrier()
        px_mc.check()
        err = (px_mc.params - p_ref).abs().max().item()
        assert err <= 2e-6, ('multicast', it, err)
        dp.assert_replicated(px_mc.params)
      w = torch.tensor([float(rank * 10 + it)], dtype=torch.float64, device='cuda')
      px.max_f64_(w, step)
      torch.cuda.synchronize()
      px.check()
      assert float(w) == (world - 1) * 10 + it
      # NCCL's summation order may differ from rank order by an ulp of the gradient sum
      err = (px.params - p_ref).abs().max().item()
      assert err <= 2e-6, (it, err)
      dp.assert_replicated(px.params)
      chunk = -(-((n - cut + world - 1) // world) // 4) * 4
      lo = cut + rank * chunk
      hi = min(n, lo + chunk)
      assert torch.allclose(m[lo:hi], m_ref[lo:hi], rtol=1e-5, atol=1e-7)   # owner keeps the moments of its shard
      assert torch.equal(m_ce[lo:hi], m[lo:hi]) and torch.equal(v_ce[lo:hi], v[lo:hi])
      assert torch.equal(shadow_ce[lo:hi], px_ce.params[lo:hi].to(torch.bfloat16))   # the owner's shard of the shadow
      step += 1
    px.close()
    px_ce.close()
    if px_mc is not None:
      px_mc.close()

    # ---- whole learner: peer exchange vs NCCL path, same seeds
    import helpers
    for precision in (0, 2):      # fp32 parity mode and the benchmarked bf16 dataflow (fused head, batched online pass, shadows)
      losses = {}
      # True = the SM-issued fused kernel; the other modes pick the halves of the fc1 + head bucket's exchange
      modes = {True: ('sm', 'ce'), 'ce': ('ce', 'ce'), 'mc': ('mcfused', 'ce'), 'mc+ce': ('mc', 'ce'), 'ce+mc': ('ce', 'mc'), False: None}
      for mode, halves in modes.items():
        if halves is not None:
          os.environ['B200RL_DP_REDUCE'], os.environ['B200RL_DP_BCAST'] = halves
        pair = helpers.make_dqn_dp_learner(rank, world, dist.group.WORLD, peer_exchange=bool(mode), precision=precision)
        if halves is not None and 'mc' in ''.join(halves) and not pair._px.multicast:   # no multicast mapping on this box
          pair._px.close()
          losses[mode] = None
          continue
        if halves is not None:
          assert pair._dp_reduce == halves[0], (pair._dp_reduce, halves)
        ls = []
        for _ in range(8):
          pair.step(fetch_loss=False)
          ls.append(float(pair.loss))
        pair.flush()                # pipelined exchange: apply the update still in flight (collective)
        torch.cuda.synchronize()
        dp.assert_replicated(pair._net.params.flat)
        if precision == 2:          # every rank's bf16 weight shadow follows the exchanged parameters
          P = pair._net.params
          assert torch.equal(P.shadow, P.flat.to(torch.bfloat16)), 'bf16 shadow out of step with the parameters'
        if pair._px is not None:
          pair._px.check()
        losses[mode] = ls
        if pair._px is not None:
          pair._px.close()
      np.testing.assert_allclose(losses[True], losses[False], rtol=2e-3 if precision == 0 else 2e-2)
      assert losses['ce'] == losses[True], (losses['ce'], losses[True])   # same sums in the same order: identical
      assert losses['ce+mc'] is None or losses['ce+mc'] == losses[True]     # only the broadcast differs: identical values
      for mode in ('mc', 'mc+ce'):                                            # in-switch sum: ulp-level differences
        if losses[mode] is not None:
          np.testing.assert_allclose(losses[mode], losses[True], rtol=2e-3 if precision == 0 else 2e-2)
    q.put((rank, 'ok'))
  except Exception as e:   # pragma: no cover
    import traceback
    q.put((rank, traceback.format_exc()))
  finally:
    dist.destroy_process_group()


def test_peer_exchange_two_gpus():
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip('needs two GPUs')
  import torch.multiprocessing as mp
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 29600 + os.getpid() % 200
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = [q.get(timeout=600) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  for rank, msg in results:
    assert msg == 'ok', f'rank {rank}: {msg}'
