import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acme_b200 import _capi
lib = _capi.load()
lib.b200rl_debug_tma_timeline.argtypes = [ctypes.c_void_p]
buf = torch.zeros(32, dtype=torch.int64, device='cuda')
ws = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
M, N, K = 256, 1024, 7744
x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda'); dy = torch.randn(M, N, device='cuda')
y = torch.empty(M, N, device='cuda'); dx = torch.empty(M, K, device='cuda'); dw = torch.empty(N, K, device='cuda')
st = _capi.current_stream()
calls = {
 'fwd': lambda: _capi.call('b200rl_linear_fwd', M, N, K, x.data_ptr(), K, w.data_ptr(), None, y.data_ptr(), N, 0, 1, ws.data_ptr(), ws.numel(), st),
 'dgrad': lambda: _capi.call('b200rl_linear_dgrad', M, N, K, dy.data_ptr(), N, w.data_ptr(), dx.data_ptr(), K, None, 0, 1, ws.data_ptr(), ws.numel(), st),
 'wgrad': lambda: _capi.call('b200rl_linear_wgrad', M, N, K, dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), None, 1, ws.data_ptr(), ws.numel(), st),
}
for name, fn in calls.items():
  for rep in range(2):
    buf.zero_(); lib.b200rl_debug_tma_timeline(buf.data_ptr())
    fn(); torch.cuda.synchronize()
    t = buf.cpu().numpy()
    print(name, rep, {i: int(v - t[0]) for i, v in enumerate(t) if v})
  lib.b200rl_debug_tma_timeline(None)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): fn()
  e1.record(); torch.cuda.synchronize()
  print(name, 'avg us', e0.elapsed_time(e1) / 20 * 1e3)
