"""Phase timestamps (globaltimer, ns) of CTA 0 of one tcgen05 GEMM launch."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acme_b200 import _capi
lib = _capi.load()
buf = torch.zeros(32, dtype=torch.int64, device='cuda')
lib.b200rl_debug_tc_timeline.argtypes = [ctypes.c_void_p]
ws = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
for (M, N, K, name) in [(256, 18, 512, 'head fwd'), (256, 1024, 7744, 'fc1 fwd')]:
  x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda'); y = torch.empty(M, N, device='cuda')
  for rep in range(3):
    buf.zero_()
    lib.b200rl_debug_tc_timeline(buf.data_ptr())
    _capi.call('b200rl_linear_fwd', M, N, K, x.data_ptr(), K, w.data_ptr(), None, y.data_ptr(), N, 0, 1, ws.data_ptr(), ws.numel(), _capi.current_stream())
    torch.cuda.synchronize()
    t = buf.cpu().numpy()
    t0 = t[0]
    print(name, rep, {i: int(v - t0) for i, v in enumerate(t) if v})
lib.b200rl_debug_tc_timeline(None)

# conv1 forward / wgrad at B=256
from acme_b200 import networks
net = networks.DQNAtariNetwork(18, precision=1)
B = 256
bufs, g = net.make_buffers(B), net.make_grad_buffers(B)
obs = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device='cuda')
P = net.params
geo = net.geom(0, B)
wsp, wsb = net.ws
for name, fn in [('conv1 fwd', lambda: _capi.call('b200rl_conv2d_fwd', obs.data_ptr(), 1, P.p('conv1.w'), P.p('conv1.b'), bufs['y1'].data_ptr(), geo, 1, 1, wsp, wsb, _capi.current_stream())),
                 ('conv1 wgrad', lambda: _capi.call('b200rl_conv2d_wgrad', obs.data_ptr(), 1, g['dy1'].data_ptr(), P.g('conv1.w'), P.g('conv1.b'), geo, 1, wsp, wsb, _capi.current_stream()))]:
  for rep in range(2):
    buf.zero_()
    lib.b200rl_debug_tc_timeline(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    t = buf.cpu().numpy()
    print(name, rep, {i: int(v - t[0]) for i, v in enumerate(t) if v})
lib.b200rl_debug_tc_timeline(None)
