"""One launch of every DQN-Atari layer kernel (B=256, tensor-core mode) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full --import-source on`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acme_b200 import _capi, networks

B = 256
net = networks.DQNAtariNetwork(18, precision=1)
bufs, g = net.make_buffers(B), net.make_grad_buffers(B)
obs = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device='cuda')
dq = torch.randn(B, 18, device='cuda')
for _ in range(3):
  net.forward(obs, bufs)
  net.backward(obs, bufs, g, dq)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.forward(obs, bufs)
net.backward(obs, bufs, g, dq)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
