"""Device time of the phases of one DQN learner step (eager, concurrent streams), CUDA events on the main stream."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from acme_b200 import _capi, adders, dqn, loggers, networks, replay, specs

spec = specs.EnvironmentSpec(specs.Array(bench.OBS_SHAPE, np.uint8), specs.DiscreteArray(18), specs.Array((), np.float32), specs.BoundedArray((), np.float32, 0., 1.))
items = 131072
table = replay.Table(replay.DEFAULT_PRIORITY_TABLE, replay.selectors.Prioritized(0.6), replay.selectors.Fifo(), max_size=items,
                     rate_limiter=replay.rate_limiters.MinSize(1), signature=adders.NStepTransitionAdder.signature(spec), max_window=3,
                     discount=0.99, slot_capacity=items + 4096, stage_slots=4096)
server = replay.Server([table])
bench.fill_replay(table, items, 3, seed=1)
prec = 1 if (len(sys.argv) < 2 or sys.argv[1] == 'bf16') else 0
net = networks.DQNAtariNetwork(18, precision=prec, seed=1)
tgt = net.clone()
ds = replay.ReplayDataset(table, 256, seed=1)
L = dqn.DQNLearner(net, tgt, 0.99, 0.2, 1e-3, 100, ds, replay_client=replay.Client(server), logger=loggers.NoOpLogger(), use_cuda_graph=False)
names = ['sample+gather', 'forwards(3)+td', 'backward', 'adam+prio+copy']
def one(ev):
  ev[0].record(); ds.sample_raw(); ev[1].record()
  # _forward_loss does forwards + td + backward; split by recording inside via monkeypatch of net.backward
  L._forward_loss_split(ev)
  L._apply(); ev[4].record()
orig_backward = net.backward
def fl_split(ev):
  def bw(*a, **k):
    ev[2].record()
    return orig_backward(*a, **k)
  net.backward = bw
  L._forward_loss()
  net.backward = orig_backward
  ev[3].record()
L._forward_loss_split = fl_split
for _ in range(5): L.step(fetch_loss=False)
tot = np.zeros(4); n = 30
for _ in range(n):
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
  one(ev); torch.cuda.synchronize()
  tot += [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(4)]
for nm, t in zip(names, tot / n): print(f'{nm:20s} {t:8.1f} us')
print('sum', (tot / n).sum())
server.stop()
