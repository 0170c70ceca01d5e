"""In-graph phase times of one DQN learner step: global-timer stamps captured between the phases (B200RL_STAMPS=1)."""
import sys, os
os.environ['B200RL_STAMPS'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from acme_b200 import _capi, adders, dqn, loggers, networks, replay, specs

world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
pg = None
if world > 1:
  import torch.distributed as dist
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
  pg = dist.group.WORLD
spec = specs.EnvironmentSpec(specs.Array(bench.OBS_SHAPE, np.uint8), specs.DiscreteArray(18), specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
items = 131072
table = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Prioritized(0.6), replay.selectors.Fifo(), max_size=items,
                     rate_limiter=replay.rate_limiters.MinSize(1), signature=adders.NStepTransitionAdder.signature(spec), max_window=3,
                     discount=0.99, slot_capacity=items + 4096, stage_slots=4096, device=rank, shard_count=world, shard_rank=rank)
server = replay.Server([table])
bench.fill_replay(table, items, 3, seed=1 + rank)
prec = {'fp32': 0, 'tf32': 1, 'bf16': 2}[sys.argv[1] if len(sys.argv) > 1 else 'bf16']
net = networks.DQNAtariNetwork(18, precision=prec, seed=1, device=rank)
tgt = net.clone()
ds = replay.ReplayDataset(table, 256, seed=1 + rank)
L = dqn.DQNLearner(net, tgt, 0.99, 0.2, 1e-3, 100, ds, replay_client=replay.Client(server), logger=loggers.NoOpLogger(), process_group=pg)
FINE = os.environ.get('B200RL_FINE', '0') == '1'
if FINE:
  marks = networks.Marks(rank)
  net.marks = tgt.marks = marks
  net.mark_prefix, tgt.mark_prefix = 'on.', 'tgt.'
ce_marks = None
if FINE and world > 1:
  import ctypes
  ce_marks = torch.zeros(8, dtype=torch.int64, device=torch.device('cuda', rank))
  _capi.load().b200rl_debug_dp_ce_marks(ctypes.c_void_p(ce_marks.data_ptr()))
names = ['k1 sample + k3 gather', 'forwards (target || batched online)', 'heads + k4 td/loss', 'backward (2 streams)', 'k7 adam', 'k2 priorities + copy + inc']
for _ in range(10): L.step(fetch_loss=False)
late = int(os.environ.get('B200RL_PHASES_AFTER', '0'))   # measure after this many extra steps (drift over long runs)
for _ in range(late): L.step(fetch_loss=False)
tot = np.zeros(6); n = 50
for _ in range(n):
  L.step(fetch_loss=False); torch.cuda.synchronize()
  t = L._stamps.cpu().numpy()
  tot += np.diff(t[:7]) / 1e3
if rank == 0:
  for nm, v in zip(names, tot / n): print(f'{nm:28s} {v:8.1f} us')
if rank == 0: print('sum', round(float((tot / n).sum()), 1), '(each stamp kernel adds ~2 us)')
if FINE and rank == 0:
  print('--- fine timeline of the last step (us since the first mark; prefix on. = online net, tgt. = target net; s1 / s2 = side streams)')
  prev = {}
  for name, t in marks.timeline():
    print(f'{t:9.1f}  {name}')
if ce_marks is not None and rank == 0:
  t = ce_marks.cpu().numpy().astype(np.float64) / 1e3
  if t[0] > 0:
    print(f'copy-engine exchange (rank 0, last step): DMA push of gradient shards {t[1]-t[0]:.1f} us, wait for the peers\' {t[2]-t[1]:.1f}, '
          f'Adam on the shard {t[3]-t[2]:.1f} | DMA push {t[5]-t[4]:.1f}, barrier {t[6]-t[5]:.1f} us')
if L._px is not None:
  import ctypes
  st = np.zeros(8, np.int64)
  _capi.load().b200rl_debug_dp_stamps(L._px._h, ctypes.c_void_p(st.ctypes.data))
  print(f'rank {rank}: last exchange kernel: wait-for-grads {(st[1]-st[0])/1e3:.1f} us, reduce+adam+broadcast {(st[2]-st[1])/1e3:.1f} us, final barrier {(st[3]-st[2])/1e3:.1f} us', flush=True)
  L.flush(); torch.cuda.synchronize()
  L._px.close()
server.stop()
if world > 1: dist.destroy_process_group()
