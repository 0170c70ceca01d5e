import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acme_b200 import _capi
lib = _capi.load()
ws = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
st = _capi.current_stream()
def t(fn, it=20):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(it): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / it * 1e3
for tma in (1, 0):
  lib.b200rl_debug_set_tma(tma)
  for (M, N, K) in [(256, 1024, 7744), (256, 1024, 8192), (256, 1024, 7680), (256, 1024, 1024), (256, 7744, 1024), (2048, 1024, 7744)]:
    x = torch.randn(M, K, device='cuda'); dy = torch.randn(M, N, device='cuda'); dw = torch.empty(N, K, device='cuda')
    us = t(lambda: _capi.call('b200rl_linear_wgrad', M, N, K, dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), None, 1, ws.data_ptr(), ws.numel(), st))
    print(f'tma={tma} wgrad M={M} N={N} K={K}: {us:8.1f} us   out {N*K*4/1e6:.1f} MB -> {N*K*4/us/1e3:.0f} GB/s')
