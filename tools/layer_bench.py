"""Per-layer device times of the DQN-Atari network at B=256 in both precisions (CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from acme_b200 import _capi, networks


def t(fn, it=20):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(it):
    fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / it * 1e3


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
  out = {}
  mode = int(os.environ.get('B200RL_STAGE_MODE', '0'))
  _capi.load().b200rl_debug_tma_stage_mode(mode)
  _capi.load().b200rl_debug_tma_dgrad_bn64(int(os.environ.get('B200RL_DGRAD_BN64', '1')))
  for prec in ((1,) if mode else (0, 1)):
    net = networks.DQNAtariNetwork(18, precision=prec)
    P = net.params
    bufs, g = net.make_buffers(B), net.make_grad_buffers(B)
    obs = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device='cuda')
    net.forward(obs, bufs)
    dq = torch.randn(B, 18, device='cuda')
    net.backward(obs, bufs, g, dq)
    ws, wsb = net.ws
    st = _capi.current_stream()
    r = {}
    x, u8 = obs.data_ptr(), 1
    for i in range(3):
      geo = net.geom(i, B)
      y = bufs[f'y{i+1}']
      r[f'conv{i+1}.fwd'] = t(lambda: _capi.call('b200rl_conv2d_fwd', x, u8, P.p(f'conv{i+1}.w'), P.p(f'conv{i+1}.b'), y.data_ptr(), geo, 1, prec, ws, wsb, st))
      dy = g[f'dy{i+1}'].data_ptr()
      r[f'conv{i+1}.wgrad'] = t(lambda: _capi.call('b200rl_conv2d_wgrad', x, u8, dy, P.g(f'conv{i+1}.w'), P.g(f'conv{i+1}.b'), geo, prec, ws, wsb, st))
      if i > 0:
        r[f'conv{i+1}.dgrad'] = t(lambda: _capi.call('b200rl_conv2d_dgrad', dy, P.p(f'conv{i+1}.w'), g[f'dy{i}'].data_ptr(), geo, bufs[f'y{i}'].data_ptr(), 1, prec, ws, wsb, st))
      x, u8 = y.data_ptr(), 0
    h, dh = bufs['h'].data_ptr(), g['dh'].data_ptr()
    F = net.flat_dim
    r['fc1.fwd'] = t(lambda: _capi.call('b200rl_linear_fwd', B, 1024, F, x, F, P.p('fc1.w'), P.p('fc1.b'), h, 1024, 1, prec, ws, wsb, st))
    r['fc1.wgrad'] = t(lambda: _capi.call('b200rl_linear_wgrad', B, 1024, F, dh, 1024, x, F, P.g('fc1.w'), P.g('fc1.b'), prec, ws, wsb, st))
    r['fc1.dgrad'] = t(lambda: _capi.call('b200rl_linear_dgrad', B, 1024, F, dh, 1024, P.p('fc1.w'), g['dy3'].data_ptr(), F, x, 1, prec, ws, wsb, st))
    r['head.fwd'] = t(lambda: _capi.call('b200rl_duelling_head_fwd', B, 18, 512, h, 1024, P.p('v2.w'), P.p('v2.b'), P.p('a2.w'),
                                         P.p('a2.b'), bufs['val'].data_ptr(), bufs['adv'].data_ptr(), bufs['q'].data_ptr(), st))
    r['head.bwd'] = t(lambda: _capi.call('b200rl_duelling_head_bwd', B, 18, 512, dq.data_ptr(), h, 1024, P.p('v2.w'), P.p('a2.w'),
                                         g['dval'].data_ptr(), g['dadv'].data_ptr(), dh, 1024, P.g('v2.w'), P.g('v2.b'),
                                         P.g('a2.w'), P.g('a2.b'), ws, wsb, st))
    r['forward'] = t(lambda: net.forward(obs, bufs))
    r['backward'] = t(lambda: net.backward(obs, bufs, g, dq))
    out['fp32' if prec == 0 else 'bf16'] = r
  if mode:
    print(json.dumps({k: round(v, 1) for k, v in out['bf16'].items()}))
    return
  for k in out['fp32']:
    print(f"{k:26s} fp32 {out['fp32'][k]:9.1f} us   bf16 {out['bf16'][k]:9.1f} us")
  json.dump(out, open('gpurun_out/layer_bench.json', 'w'), indent=1)


main()
