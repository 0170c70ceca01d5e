"""Life of ONE CTA of the bf16 conv data-gradient kernel (global-timer stamps inside the kernel): where do the ~10 us of
a 4-k-block tile go?  Runs conv2's and conv3's dgrad at B = 256 standalone and prints the stamps of the middle CTA."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_b200 import _capi, networks

lib = _capi.load()
buf = torch.zeros(16, dtype=torch.int64, device='cuda')
lib.b200rl_debug_h_timeline(ctypes.c_void_p(buf.data_ptr()))
names = ['cta start', 'set-up done (barriers, TMEM alloc, sync)', 'first operands landed', 'last operands landed',
         'accumulator ready (epilogue woke)', 'TMEM -> smem slab done', 'global stores issued', 'cta exit']
net = networks.DQNAtariNetwork(18, precision=2, seed=0)
B = 256
for i in (2, 1):
  k, s, ci, co, h, oh, pad = net.convs[i]
  g = net.geom(i, B)
  dy = torch.randn(B, oh, oh, co, device='cuda').to(torch.bfloat16)
  w = torch.randn(co, k, k, ci, device='cuda').to(torch.bfloat16)
  y = torch.randn(B, h, h, ci, device='cuda').to(torch.bfloat16)
  dx = torch.empty(B, h, h, ci, device='cuda', dtype=torch.bfloat16)
  call = lambda: _capi.call('b200rl_conv2d_dgrad_bf16', dy.data_ptr(), w.data_ptr(), dx.data_ptr(), 1, g, y.data_ptr(), 1, 1,
                            _capi.current_stream())
  for _ in range(3):
    call()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20):
    call()
  e1.record(); torch.cuda.synchronize()
  t = buf.cpu().numpy()
  print(f'conv{i + 1} dgrad B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (warm, back to back)')
  for j, nm in enumerate(names):
    print(f'   {(t[j] - t[0]) / 1e3:7.2f} us  {nm}')
