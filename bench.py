#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: DQN learner updates/s (and PER samples/s) on
the Atari-shaped workload of BASELINE.json configs[1]:
  uint8 84x84x4 observations, prioritized replay of 1,000,000 steps (alpha 0.6), batch 256, n-step 3,
  DQNAtariNetwork(18), Adam 1e-3, IS exponent 0.2, target period 100, synthetic data, random init.

One "step" = one full learner update (K1 sample -> K3 gather/n-step -> 3 forwards -> K4 -> backward ->
K7 Adam -> K2 priority write-back -> target copy) over one batch per rank.

  python bench.py [--gpus N --steps K --warmup W]          our arm (one rank per GPU under torchrun)
  python bench.py --impl reference [...]                  the reference's CPU path (oracle port), host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'learner_updates_per_sec_dqn_atari_b256'
UNIT = 'updates/s (one update = 256 transitions; N ranks do N x 256 per data-parallel step)'
OBS_SHAPE = (84, 84, 4)
NUM_ACTIONS = 18
STEP_FLOPS = 5 * 2 * 19_977_728 * 256     # 3 fwd + bwd(2x) of 19.98 MMAC/sample at B=256


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'], src='measured')
  return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


# ------------------------------------------------------------------------------ synthetic workload
class SynthStream:
  """Host-side scalars of a synthetic actor stream: actions U{0..17}, rewards in
  {-1,0,1} w.p. {.05,.9,.05}, env discount 1 except 0 w.p. 1/1000, geometric episode lengths (mean
  ~1000 steps).  Element i is one observation slot; first/last mark episode boundaries and stay
  consistent across chunk boundaries."""

  def __init__(self, seed, mean_episode=1000):
    self.rng = np.random.default_rng(seed)
    self.mean = mean_episode
    self.left = 0          # slots left in the current episode (0 = next slot starts a new one)

  def chunk(self, n):
    rng = self.rng
    act = rng.integers(0, NUM_ACTIONS, n).astype(np.int32)
    rew = rng.choice(np.array([-1., 0., 1.], np.float32), n, p=[.05, .9, .05]).astype(np.float32)
    disc = np.where(rng.random(n) < 1e-3, 0., 1.).astype(np.float32)
    first, last = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    i = 0
    while i < n:
      if self.left == 0:
        self.left = int(rng.geometric(1.0 / self.mean)) + 1      # T steps -> T+1 slots, T >= 1
        first[i] = 1
      take = min(self.left, n - i)
      self.left -= take
      i += take
      if self.left == 0:
        last[i - 1] = 1
    return act, rew, disc, first, last


def fill_replay(table, steps, n_step, seed, chunk=32768):
  """Fills the HBM ring with `steps` synthetic timesteps (observations generated on the device:
  setup only, not timed) and then sets priorities to |N(0,1)| like SURVEY §8d."""
  import ctypes as C
  import torch
  from acme_b200 import _capi
  stream = SynthStream(seed)
  wid = C.c_int32()
  _capi.call('b200rl_writer_open', table.handle, C.byref(wid))
  gen = torch.Generator(device='cuda')
  gen.manual_seed(seed)
  done = 0
  while done < steps:
    n = min(chunk, steps - done)
    act, rew, disc, first, last = stream.chunk(n)
    obs = torch.randint(0, 256, (n,) + OBS_SHAPE, dtype=torch.uint8, device='cuda', generator=gen)
    _capi.call('b200rl_writer_append_stream', table.handle, wid.value, n, obs.data_ptr(), 1, act.ctypes.data,
               rew.ctypes.data, disc.ctypes.data, first.ctypes.data, last.ctypes.data, n_step, 1.0,
               _capi.current_stream())
    torch.cuda.synchronize()
    del obs
    done += n
  info = table.info()
  size, tail = info['size'], info['tail_key']
  keys = (torch.arange(size, dtype=torch.int64, device='cuda') + tail).view(torch.uint64)
  pr = torch.randn(size, device='cuda', generator=gen).abs_()
  table.update_priorities_device(keys, pr)
  torch.cuda.synchronize()
  return info


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
  """nvidia-smi sampler (the recipe's clocks line).  It is started before the warm-up and samples every 50 ms;
  `mark()` brackets the timed region so only samples taken inside it are summarised (if the region is shorter
  than the sampling period, the samples nearest to it are used)."""
  Q = ('timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index=0):
    self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
    self.p = None
    self.t0 = self.t1 = None
    try:
      self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '50',
                                 '-i', str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
    except Exception:  # noqa: BLE001
      self.p = None

  def mark(self):
    if self.t0 is None:
      self.t0 = time.time()
    else:
      self.t1 = time.time()

  def stop(self):
    import datetime
    if self.p is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.12)
    self.p.terminate()
    self.p.wait()
    self.f.flush()
    rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
    os.unlink(self.f.name)
    samples = []
    for r in rows:
      try:
        ts = datetime.datetime.strptime(r[0].strip(), '%Y/%m/%d %H:%M:%S.%f').timestamp()
        samples.append((ts, float(r[2]), float(r[3]), [v.strip().lower().startswith('active') for v in r[6:10]]))
      except Exception:  # noqa: BLE001
        continue
    if not samples:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
    t0, t1 = self.t0 or samples[0][0], self.t1 or samples[-1][0]
    inside = [x for x in samples if t0 - 0.05 <= x[0] <= t1 + 0.05]
    if not inside:   # region shorter than the sampling period: take the two samples nearest to it
      inside = sorted(samples, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[:2]
    names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
    reasons = sorted({n for x in inside for n, on in zip(names, x[3]) if on})
    return {'sm_mhz': float(np.median([x[1] for x in inside])), 'sm_max_mhz': float(max(x[2] for x in inside)),
            'reasons': reasons, 'samples': len(inside)}


# ------------------------------------------------------------------------------ CPU baseline / reference arm
class CpuReferencePath:
  """The reference's CPU path restated: binary f64 sum-tree sample + update
  (Reverb Prioritized-like), NumPy gather + n-step, PyTorch-CPU DQN learner step (oracle.learner).
  Bounded sample of the workload: same tree size (1M priorities), same batch/network, observation
  ring of `ring` steps (28 KB/step makes a 1M-step host ring impractical)."""

  def __init__(self, items=1_000_000, ring=16384, batch=256, n_step=3, seed=0, threads=None):
    import torch
    from oracle import learner as olearner
    from oracle import nets as onets
    from oracle import sumtree
    self.threads = threads or os.cpu_count()
    torch.set_num_threads(self.threads)
    rng = np.random.default_rng(seed)
    self.rng, self.B, self.n, self.ring, self.items = rng, batch, n_step, ring, items
    self.tree = sumtree.BinarySumTreeF64(items)
    self.tree.build(np.abs(rng.standard_normal(items))**0.6)
    self.obs = rng.integers(0, 256, (ring,) + OBS_SHAPE, dtype=np.uint8)
    self.act, self.rew, self.disc, _, _ = SynthStream(seed).chunk(ring)
    net, tgt = onets.DQNAtariNetwork(NUM_ACTIONS, seed=seed), onets.DQNAtariNetwork(NUM_ACTIONS, seed=seed)
    tgt.copy_from(net)
    self.learner = olearner.DQNOracleLearner(net, tgt, 0.99, 0.2, 1e-3, 100)
    self.g = np.float32(0.99)

  def step(self):
    u = self.rng.random(self.B)
    idx, prob = self.tree.sample(u)                       # 256 root-to-leaf walks
    s = idx % (self.ring - self.n - 1)
    R = self.rew[s].copy()
    D = self.disc[s].copy()
    for j in range(1, self.n):                            # transition.py:135-145, vectorised over the batch
      D = D * self.g
      R = R + self.rew[s + j] * D
      D = D * self.disc[s + j]
    out = self.learner.step(self.obs[s], self.act[s], R, D, self.obs[s + self.n], prob.astype(np.float32))
    self.tree.set(idx, out['priority']**0.6)              # update_priorities
    return float(out['loss'])


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  t0 = time.time()
  ref = CpuReferencePath(seed=1234)
  for _ in range(args.warmup):
    ref.step()
  t1 = time.time()
  for _ in range(args.steps):
    ref.step()
  dt = (time.time() - t1) / max(args.steps, 1)
  cpu = open('/proc/cpuinfo').read()
  model = next((l.split(':', 1)[1].strip() for l in cpu.splitlines() if l.startswith('model name')), 'unknown')
  value = 1.0 / dt
  sample = (f'{args.steps} full learner updates (B=256, DQNAtariNetwork(18), binary f64 sum-tree over 1M priorities, '
            f'{ref.ring}-step uint8 observation ring), PyTorch-CPU + NumPy, {ref.threads} threads, {model}')
  print(json.dumps({
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'DQN Atari-shaped 84x84x4 uint8, PER 1M items, batch 256, n=3 (BASELINE configs[1])',
                 'note': 'reference arm = CPU restatement (oracle port); Acme TF/JAX + Reverb are not installable offline'},
      'cpu_baseline': {'value': value, 'unit': 'updates/s', 'cores': ref.threads, 'kind': 'port', 'sample': sample},
      'e2e': {'value': value, 'unit': 'updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'setup_s': t1 - t0,
  }))


# ------------------------------------------------------------------------------ our arm
def time_stage(fn, iters, torch):
  """Average device time of fn() over `iters` calls, CUDA events on the current stream."""
  fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e-3


def run_ours(args):
  import torch
  import torch.distributed as dist
  from acme_b200 import _capi, adders, dm_env, dqn, loggers, networks, replay, specs

  rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  _capi.require_device(local)
  pg = None
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    pg = dist.group.WORLD
  P = peaks()
  B, n_step, items = 256, 3, args.items
  precision = {'fp32': _capi.PRECISION_FP32, 'tf32': _capi.PRECISION_TC, 'tc': _capi.PRECISION_TC,
               'bf16': _capi.PRECISION_BF16FLOW}[args.precision]

  spec = specs.EnvironmentSpec(specs.Array(OBS_SHAPE, np.uint8), specs.DiscreteArray(NUM_ACTIONS),
                               specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
  table = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Prioritized(0.6), replay.selectors.Fifo(),
                       max_size=items, rate_limiter=replay.rate_limiters.MinSize(1),
                       signature=adders.NStepTransitionAdder.signature(spec), max_window=n_step, discount=0.99,
                       device=local, slot_capacity=items + 4096, shard_count=world, shard_rank=rank, stage_slots=4096)
  server = replay.Server([table])
  t_setup = time.time()
  info = fill_replay(table, items, n_step, seed=1234 + rank)
  net = networks.DQNAtariNetwork(NUM_ACTIONS, device=local, precision=precision, seed=1234)   # replicated init
  tgt = net.clone()
  ds = replay.ReplayDataset(table, B, seed=1234 + rank)
  client = replay.Client(server)
  learner = dqn.DQNLearner(net, tgt, 0.99, 0.2, 1e-3, 100, ds, replay_client=client, logger=loggers.NoOpLogger(),
                           process_group=pg, use_cuda_graph=not args.no_graph)
  adder = adders.NStepTransitionAdder(client, n_step=n_step, discount=0.99)
  t_setup = time.time() - t_setup
  lib = _capi.load()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident timing (value)
  clocks = ClockSampler(local)
  for _ in range(max(args.warmup, 3)):
    learner.step(fetch_loss=False)
  barrier()
  launches0 = lib.b200rl_launch_count()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  clocks.mark()
  if args.profile:
    torch.cuda.profiler.start()     # ncu --profile-from-start off: only the timed learner steps are captured
  e0.record()
  for _ in range(args.steps):
    learner.step(fetch_loss=False)
  learner.flush()                   # pipelined data-parallel exchange: the last update is applied inside the timed region
  e1.record()
  barrier()
  if args.profile:
    torch.cuda.profiler.stop()
  clocks.mark()
  dev_s = e0.elapsed_time(e1) * 1e-3
  clk = clocks.stop()
  if world > 1:
    tt = torch.tensor([dev_s], device='cuda')
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_s = float(tt)
  ms_per_step = dev_s / args.steps * 1e3
  value = world * args.steps / dev_s

  if args.profile:
    if rank == 0:
      print(json.dumps({'metric': METRIC, 'value': value, 'ms_per_step': ms_per_step, 'profile_run': True}))
    server.stop()
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- end to end through the public API with host buffers: the reference loop's cadence is 8
  #      environment steps inserted per learner step (batch 256 / samples_per_insert 32, dqn/agent.py:158-162)
  rng = np.random.default_rng(99 + rank)
  host_obs = [rng.integers(0, 256, OBS_SHAPE, dtype=np.uint8) for _ in range(64)]
  inserts_per_step = 8
  adder.add_first(dm_env.restart(host_obs[0]))

  def e2e_step(i):
    for j in range(inserts_per_step):
      o = host_obs[(i * inserts_per_step + j + 1) % 64]
      adder.add(np.int32(j % NUM_ACTIONS), dm_env.transition(np.float32(0.), o, np.float32(1.)))
    # flush (H2D of the staged steps) + update + D2H of the loss into pinned memory; the loss of update i is consumed
    # (logged) when update i+1 has been issued, so the host-side inserts overlap the device step
    learner.step(fetch_loss='async')

  for i in range(3):
    e2e_step(i)
  barrier()
  t0 = time.perf_counter()
  e2e_steps = max(args.steps // 2, 10)
  for i in range(e2e_steps):
    e2e_step(i + 3)
  learner.flush()
  last_loss = learner.drain()            # the final update's loss, read inside the timed region
  assert last_loss is not None and np.isfinite(last_loss)
  barrier()
  e2e_s = time.perf_counter() - t0
  if world > 1:
    tt = torch.tensor([e2e_s], device='cuda')
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s = float(tt)
  e2e_value = world * e2e_steps / e2e_s
  h2d = inserts_per_step * (int(np.prod(OBS_SHAPE)) + 4 + 4 + 4 + 4) + inserts_per_step * 24 + 16
  d2h = 4

  # ---- per-stage device times (eager launches, CUDA events) and the roofline of the dominant kernel
  stages = {}
  if rank == 0:
    it = 20
    u = torch.rand(B, device='cuda')
    stages['k1_sample'] = time_stage(lambda: table.sample_into(u, ds.idx, ds.keys, ds.prob, True), it, torch)
    stages['k3_gather_nstep'] = time_stage(lambda: table.gather_into(ds.idx, ds.o_tm1, ds.a_tm1, ds.R, ds.D, ds.o_t), it, torch)
    # the network group exactly as the step runs it: target pass beside the batched online pass, then the fused
    # head + K4 kernel and the backward (weight gradients beside data gradients), recorded into ONE graph and replayed,
    # so the figure is the in-step time of the group, not a sum of serialised eager launches
    def join_tail():     # B200RL_SPLIT_ADAM=1 forks the fc1 + head optimizer update inside the backward: rejoin for the capture
      if getattr(learner, '_tail_done', None) is not None:
        torch.cuda.current_stream().wait_event(learner._tail_done)
        learner._tail_done = None
    g_fwd = learner._capture(learner._forwards)
    g_bwd = learner._capture(lambda: (learner._loss_backward(), join_tail()))
    g_net = learner._capture(lambda: (learner._forwards(), learner._loss_backward(), join_tail()))
    stages['k6_forwards_in_graph'] = time_stage(g_fwd.replay, it, torch)
    stages['k4_k6_loss_backward_in_graph'] = time_stage(g_bwd.replay, it, torch)
    stages['k6_network_group_in_graph'] = time_stage(g_net.replay, it, torch)
    Pn = net.params
    scratch_p, scratch_m, scratch_v = torch.zeros_like(Pn.flat), torch.zeros_like(Pn.flat), torch.zeros_like(Pn.flat)
    stages['k7_adam'] = time_stage(lambda: _capi.call(
        'b200rl_adam', Pn.size, scratch_p.data_ptr(), Pn.grad.data_ptr(), scratch_m.data_ptr(), scratch_v.data_ptr(),
        learner._num_steps.data_ptr(), 1e-3, 0.9, 0.999, 1e-8, 0, None, None, _capi.current_stream()), it, torch)
    stages['k2_update_priorities'] = time_stage(lambda: table.update_priorities_device(ds.keys, learner.priority), it, torch)
    del scratch_p, scratch_m, scratch_v
  net_s = stages.get('k6_network_group_in_graph', 0)

  # ---- PER sampling throughput (sample + gather), batch 256 and a large-batch sweep point
  per = {}
  if rank == 0:
    per['samples_per_sec_b256_sample_gather'] = B / (stages['k1_sample'] + stages['k3_gather_nstep'])
    Bl = 1 << 16
    ul = torch.rand(Bl, device='cuda')
    il = torch.empty(Bl, dtype=torch.int64, device='cuda')
    kl = torch.empty(Bl, dtype=torch.uint64, device='cuda')
    pl = torch.empty(Bl, device='cuda')
    t_l = time_stage(lambda: table.sample_into(ul, il, kl, pl, True), 20, torch)
    L, F, S = table.tree_levels()
    per['sample_only_b65536'] = {'samples_per_sec': Bl / t_l, 'us': t_l * 1e6,
                                 'alg_bytes_per_sample': (L - min(S + 1, L)) * F * 4 + 24, 'levels': L, 'fanout': F}

  # ---- CPU baseline beside it (rank 0, N=1 only, bounded)
  cpu_baseline = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    ref = CpuReferencePath(seed=1234, ring=8192)
    ref.step()
    t0 = time.time()
    n_cpu = 0
    while n_cpu < 3 or (time.time() - t0 < 12 and n_cpu < 30):
      ref.step()
      n_cpu += 1
    cdt = (time.time() - t0) / n_cpu
    cpu_baseline = {'value': 1.0 / cdt, 'unit': 'updates/s', 'cores': ref.threads, 'kind': 'port',
                    'sample': f'{n_cpu} full learner updates of the same workload on the host (oracle port: binary f64 '
                              f'sum-tree over 1M priorities, 8192-step uint8 ring, PyTorch-CPU DQNAtariNetwork(18) B=256)'}

  if rank == 0:
    gpu_launches = int(learner.kernel_launches_per_step or 0) * args.steps
    if not gpu_launches:
      gpu_launches = int(lib.b200rl_launch_count() - launches0)
    kname = {0: 'k6_network_gemm_conv (3 forwards + 1 backward, fp32 SIMT FFMA)',
             1: 'k6_network_gemm_conv (3 forwards + 1 backward; tcgen05 kind::tf32, TMA-fed, fp32 tensors)',
             2: 'k6_network_gemm_conv (3 forwards + 1 backward; tcgen05 kind::f16 on bf16 activations / weight shadows / '
                'gradients, TMA-fed: im2col for the convolutions incl. conv1 on uint8 frames via an integer-valued bf16 row image)'}[precision]
    achieved = STEP_FLOPS / net_s / 1e12 if net_s > 0 else None
    traffic = None    # dram bytes of the same kernel group from the committed `ncu --set full` capture (tensor-core mode)
    prof_name = {1: 'ncu_full_r01_tc_layers_summary.json', 2: 'ncu_full_r02_bf16_step_summary.json'}.get(precision)
    prof = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', prof_name or 'none')
    traffic_source = None
    if os.path.exists(prof):
      traffic = json.load(open(prof)).get('step_group_dram_bytes')
      traffic_source = f'profiles/{prof_name} (committed ncu --set full capture of the same kernels; not measured in this run)'
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {0: 'f32', 1: 'tf32 tensor-core operands, f32 accumulate, f32 master weights and optimizer',
                  2: 'bf16 tensor-core operands (activations, weight shadows, back-propagated gradients), f32 accumulate, '
                     'f32 master weights, weight gradients and optimizer'}[precision],
        'data': 'synthetic',
        'config': {'workload': 'DQN Atari-shaped 84x84x4 uint8, PER 1M items, batch 256, n=3 (BASELINE configs[1])',
                   'items_per_rank': info['size'], 'batch_per_rank': B, 'n_step': n_step, 'alpha': 0.6, 'beta': 0.2,
                   'network': 'DQNAtariNetwork(18), 8,018,611 params', 'optimizer': 'Adam 1e-3',
                   'parallelism': (f'dp{world}: per-rank replay shard; gradient mean + Adam + parameter broadcast fused over '
                                   'NVLink peer memory (NCCL for set-up only)') if world > 1 else 'single GPU',
                   'cuda_graph': not args.no_graph,
                   'l2_policy': 'inputs larger than L2: 28 GB ring sampled at random + 160 MB of params/moments/grads per step'},
        'clocks': clk,
        'e2e': {'value': e2e_value, 'unit': 'updates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'what': f'{inserts_per_step} adder.add() calls from host numpy (pinned staging -> H2D) + learner.step() '
                        "with every update's loss copied to pinned host memory and consumed one update later (fetch_loss='async'; the last one "
                        'inside the timed region), per update (the reference loop cadence)', 'steps': e2e_steps},
        'gpu_launches': gpu_launches,
        'roofline': {'kernel': kname, 'bound': 'tensor', 'achieved': achieved, 'peak': P['tf_sustained'], 'unit': 'TFLOP/s',
                     'frac': (achieved / P['tf_sustained']) if achieved else None, 'traffic': traffic,
                     'peak_source': f"{P['src']} (sustained bf16 dense peak: the group is timed inside a long run; "
                                    'tf32 runs at half that rate)',
                     'traffic_source': traffic_source, 'alg_flops_per_launch_group': STEP_FLOPS, 'group_seconds': net_s,
                     'group_timing': 'CUDA events around replays of the captured network group (forwards + head/K4 + backward)'},
        'cpu_baseline': cpu_baseline,
        'stages_us': {k: v * 1e6 for k, v in stages.items()},
        'hbm_kernels': {
            'k3_gather_nstep': {'alg_bytes': B * 112_944, 'GBps': B * 112_944 / stages['k3_gather_nstep'] / 1e9,
                                'frac_of_hbm_peak': B * 112_944 / stages['k3_gather_nstep'] / 1e9 / P['hbm']},
            'k7_adam': {'alg_bytes': net.params.size * 28, 'GBps': net.params.size * 28 / stages['k7_adam'] / 1e9,
                        'frac_of_hbm_peak': net.params.size * 28 / stages['k7_adam'] / 1e9 / P['hbm']},
        },
        'per': per,
        'setup_s': t_setup,
    }
    print(json.dumps(out))
  if world > 1:   # replicas must not have diverged (identical Adam on identical averaged gradients)
    learner._dp.assert_replicated(net.params.flat, 'online parameters')
  server.stop()
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=2000)
  ap.add_argument('--warmup', type=int, default=50)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--precision', default=os.environ.get('B200RL_PRECISION', 'bf16'), choices=['fp32', 'tf32', 'tc', 'bf16'],
                  help="'bf16' (default): bf16 dataflow on tcgen05 (kind::f16); 'tf32' ('tc'): fp32 tensors with tf32 operands; "
                       "both with the stated tolerances of tests/test_gpu_learner.py; 'fp32': SIMT FFMA, 1e-5 parity with the oracle")
  ap.add_argument('--items', type=int, default=1_000_000)
  ap.add_argument('--no-graph', action='store_true')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--profile', action='store_true', help='only warm-up + timed steps (for ncu captures)')
  args = ap.parse_args()
  if args.impl == 'reference':
    if args.steps == 100 and args.warmup == 10:      # defaults sized for the GPU arm; bound the CPU run
      args.steps, args.warmup = 10, 2
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
