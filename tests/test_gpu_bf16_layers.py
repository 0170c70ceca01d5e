"""bf16 dataflow kernels (gemm_bf16.cu) through the C ABI against PyTorch-CPU fp32 on the SAME bf16-rounded operands: the
only differences left are the fp32 summation order and the rounding of a bf16 output (2^-9 relative), so the bounds are
tight.  Shapes: the DQN-Atari layers at small batch plus the benchmarked batch for the dense layer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_KEEP = []


def dev(x, dtype=None):
  import torch
  t = torch.as_tensor(np.ascontiguousarray(x)).cuda()
  if dtype is not None:
    t = t.to(dtype)
  _KEEP.append(t)
  return t


@pytest.fixture(autouse=True)
def _release():
  yield
  import torch
  torch.cuda.synchronize()
  _KEEP.clear()


def bf(x):
  """fp32 array rounded to bf16 (as fp32)."""
  import torch
  return torch.as_tensor(np.ascontiguousarray(x, np.float32)).to(torch.bfloat16).float().numpy()


def close(got, want, tol, name=''):
  got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
  scale = max(float(np.abs(want).max()), 1e-30)
  err = float(np.abs(got - want).max()) / scale
  assert err <= tol, f'{name}: max error {err:.3e} of scale exceeds {tol:.1e}'


def ws():
  import torch
  t = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
  _KEEP.append(t)
  return t.data_ptr(), t.numel()


OUT_BF16_TOL, OUT_F32_TOL = 6e-3, 2e-4


@pytest.mark.parametrize('M,N,K', [(256, 1024, 7744), (512, 1024, 7744), (64, 128, 192), (100, 64, 64)])
def test_linear_bf16(M, N, K):
  import torch
  from acme_b200 import _capi
  rng = np.random.default_rng(M + N + K)
  x, w = bf(rng.standard_normal((M, K))), bf(rng.standard_normal((N, K)) / np.sqrt(K))
  b = rng.standard_normal(N).astype(np.float32)
  dy = bf(rng.standard_normal((M, N)))
  xd, wd, dyd = dev(x, torch.bfloat16), dev(w, torch.bfloat16), dev(dy, torch.bfloat16)
  wsp, wsb = ws()
  st = _capi.current_stream()
  # forward, fp32 and bf16 outputs
  y = np.maximum(x @ w.T + b, 0)
  y32 = torch.empty(M, N, device='cuda')
  _capi.call('b200rl_linear_fwd_bf16', M, N, K, xd.data_ptr(), K, wd.data_ptr(), dev(b).data_ptr(), y32.data_ptr(), N, 0,
             _capi.ACT_RELU, wsp, wsb, st)
  close(y32.cpu().numpy(), y, OUT_F32_TOL, 'linear fwd (fp32 out)')
  y16 = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
  _capi.call('b200rl_linear_fwd_bf16', M, N, K, xd.data_ptr(), K, wd.data_ptr(), dev(b).data_ptr(), y16.data_ptr(), N, 1,
             _capi.ACT_RELU, wsp, wsb, st)
  close(y16.float().cpu().numpy(), y, OUT_BF16_TOL, 'linear fwd (bf16 out)')
  # data gradient with the ReLU mask of a bf16 producer output
  mask = bf(rng.standard_normal((M, K)))
  dx = (dy @ w) * (mask > 0)
  dx16 = torch.empty(M, K, device='cuda', dtype=torch.bfloat16)
  _capi.call('b200rl_linear_dgrad_bf16', M, N, K, dyd.data_ptr(), N, wd.data_ptr(), dx16.data_ptr(), K, 1,
             dev(mask, torch.bfloat16).data_ptr(), 1, _capi.ACT_RELU, wsp, wsb, st)
  close(dx16.float().cpu().numpy(), dx, OUT_BF16_TOL, 'linear dgrad')
  # weight gradient + bias gradient (fp32 outputs)
  dw, db = dy.T @ x, dy.sum(0)
  dwd, dbd = torch.empty(N, K, device='cuda'), torch.empty(N, device='cuda')
  _capi.call('b200rl_linear_wgrad_bf16', M, N, K, dyd.data_ptr(), N, xd.data_ptr(), K, dwd.data_ptr(), dbd.data_ptr(), wsp, wsb, st)
  close(dwd.cpu().numpy(), dw, OUT_F32_TOL, 'linear wgrad')
  close(dbd.cpu().numpy(), db, OUT_F32_TOL, 'linear bias grad')


def _geom(B, H, C, k, s, Cout):
  from acme_b200 import _capi, networks
  oh, pad = networks.tf_same_pad(H, k, s)
  return _capi.ConvGeom(B=B, H=H, W=H, C=C, kh=k, kw=k, stride=s, pad_top=pad, pad_left=pad, OH=oh, OW=oh, Cout=Cout), oh, pad


def _torch_conv(x_nhwc, w_ohwi, b, k, s, H, oh, pad):
  """TF-SAME convolution in PyTorch-CPU fp32 (explicit asymmetric padding); returns (y NHWC, autograd handles)."""
  import torch
  import torch.nn.functional as F
  xt = torch.tensor(x_nhwc).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
  wt = torch.tensor(w_ohwi).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
  bt = torch.tensor(b).requires_grad_(True)
  total = max((oh - 1) * s + k - H, 0)
  xp = F.pad(xt, (pad, total - pad, pad, total - pad))
  y = F.conv2d(xp, wt, bt, stride=s)
  return y, xt, wt, bt


@pytest.mark.parametrize('B,H,C,k,s,Cout', [(6, 21, 32, 4, 2, 64), (6, 11, 64, 3, 1, 64), (40, 21, 32, 4, 2, 64),
                                             (40, 11, 64, 3, 1, 64)])
def test_conv_bf16(B, H, C, k, s, Cout):
  """conv2 / conv3 of the Atari torso (atari.py:45-48): forward + ReLU, data gradient with the producer's ReLU mask,
  weight and bias gradients."""
  import torch
  from acme_b200 import _capi
  rng = np.random.default_rng(B + H + C)
  g, oh, pad = _geom(B, H, C, k, s, Cout)
  x = bf(np.maximum(rng.standard_normal((B, H, H, C)), 0))
  w = bf(rng.standard_normal((Cout, k, k, C)) / np.sqrt(k * k * C))
  b = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
  y, xt, wt, bt = _torch_conv(x, w, b, k, s, H, oh, pad)
  yr = torch.relu(y)
  dy = bf(rng.standard_normal((B, oh, oh, Cout)) * (yr.permute(0, 2, 3, 1).detach().numpy() > 0))   # d(pre-activation)
  y.backward(torch.tensor(dy).permute(0, 3, 1, 2))
  wsp, wsb = ws()
  st = _capi.current_stream()
  xd, wd, dyd = dev(x, torch.bfloat16), dev(w, torch.bfloat16), dev(dy, torch.bfloat16)
  yd = torch.empty(B, oh, oh, Cout, device='cuda', dtype=torch.bfloat16)
  _capi.call('b200rl_conv2d_fwd_bf16', xd.data_ptr(), 0, wd.data_ptr(), dev(b).data_ptr(), yd.data_ptr(), 1, g, _capi.ACT_RELU,
             wsp, wsb, st)
  close(yd.float().cpu().numpy(), yr.permute(0, 2, 3, 1).detach().numpy(), OUT_BF16_TOL, 'conv fwd')
  dxd = torch.empty(B, H, H, C, device='cuda', dtype=torch.bfloat16)
  _capi.call('b200rl_conv2d_dgrad_bf16', dyd.data_ptr(), wd.data_ptr(), dxd.data_ptr(), 1, g, xd.data_ptr(), 1, _capi.ACT_RELU, st)
  close(dxd.float().cpu().numpy(), xt.grad.permute(0, 2, 3, 1).numpy() * (x > 0), OUT_BF16_TOL, 'conv dgrad')
  dwd, dbd = torch.empty(Cout, k, k, C, device='cuda'), torch.empty(Cout, device='cuda')
  _capi.call('b200rl_conv2d_wgrad_bf16', xd.data_ptr(), 0, dyd.data_ptr(), dwd.data_ptr(), dbd.data_ptr(), g, wsp, wsb, st)
  close(dwd.cpu().numpy(), wt.grad.permute(0, 2, 3, 1).numpy(), OUT_F32_TOL, 'conv wgrad')
  close(dbd.cpu().numpy(), bt.grad.numpy(), OUT_F32_TOL, 'conv bias grad')


@pytest.mark.parametrize('B', [5, 33])
def test_conv1_rows_bf16(B):
  """First layer on uint8 frames (atari.py:44, atari_wrapper.py:303-304): integer-valued bf16 row image, 1/255 applied to
  the accumulator; forward and weight gradient."""
  import ctypes
  import torch
  from acme_b200 import _capi
  rng = np.random.default_rng(B)
  H, C, k, s, Cout = 84, 4, 8, 4, 32
  g, oh, pad = _geom(B, H, C, k, s, Cout)
  frames = rng.integers(0, 256, (B, H, H, C), dtype=np.uint8)
  w = bf(rng.standard_normal((Cout, k, k, C)) / np.sqrt(k * k * C))
  b = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
  x = (frames.astype(np.float32) / np.float32(255)).astype(np.float32)
  y, xt, wt, bt = _torch_conv(x, w, b, k, s, H, oh, pad)
  yr = torch.relu(y)
  dy = bf(rng.standard_normal((B, oh, oh, Cout)) * (yr.permute(0, 2, 3, 1).detach().numpy() > 0))
  y.backward(torch.tensor(dy).permute(0, 3, 1, 2))
  nbytes = int(_capi.load().b200rl_conv2d_rows_bf16_bytes(ctypes.byref(g)))
  assert nbytes > 0
  rows = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
  st = _capi.current_stream()
  _capi.call('b200rl_conv2d_rows_bf16_from_u8', dev(frames).data_ptr(), g, rows.data_ptr(), nbytes, st)
  wsp, wsb = ws()
  yd = torch.empty(B, oh, oh, Cout, device='cuda', dtype=torch.bfloat16)
  _capi.call('b200rl_conv2d_fwd_bf16', rows.data_ptr(), 1, dev(w, torch.bfloat16).data_ptr(), dev(b).data_ptr(), yd.data_ptr(), 1,
             g, _capi.ACT_RELU, wsp, wsb, st)
  close(yd.float().cpu().numpy(), yr.permute(0, 2, 3, 1).detach().numpy(), OUT_BF16_TOL, 'conv1 fwd')
  dwd, dbd = torch.empty(Cout, k, k, C, device='cuda'), torch.empty(Cout, device='cuda')
  _capi.call('b200rl_conv2d_wgrad_bf16', rows.data_ptr(), 1, dev(dy, torch.bfloat16).data_ptr(), dwd.data_ptr(), dbd.data_ptr(),
             g, wsp, wsb, st)
  close(dwd.cpu().numpy(), wt.grad.permute(0, 2, 3, 1).numpy(), OUT_F32_TOL, 'conv1 wgrad')
  close(dbd.cpu().numpy(), bt.grad.numpy(), OUT_F32_TOL, 'conv1 bias grad')


def test_fused_head_td_matches_unfused_kernels():
  """b200rl_dqn_head_td == duelling_head_fwd x3 + dqn_td + duelling_head_bwd (dh) bit for bit."""
  import torch
  from acme_b200 import _capi
  rng = np.random.default_rng(8)
  B, A, H = 256, 18, 512
  hs = [np.maximum(rng.standard_normal((B, 2 * H)), 0).astype(np.float32) for _ in range(3)]     # tm1, sel, tgt
  def head():
    return (rng.standard_normal(H).astype(np.float32) * 0.05, rng.standard_normal(1).astype(np.float32),
            rng.standard_normal((A, H)).astype(np.float32) * 0.05, rng.standard_normal(A).astype(np.float32))
  on, tg = head(), head()
  a = rng.integers(0, A, B).astype(np.int32)
  R = rng.choice([-1., 0., 1., 2.5], B).astype(np.float32)
  D = rng.choice([0., 0.9801, 1.0], B).astype(np.float32)
  prob = rng.uniform(1e-7, 1e-3, B).astype(np.float32)
  st = _capi.current_stream()
  f = lambda *s: torch.empty(s, device='cuda')
  hd = [dev(h) for h in hs]
  ond, tgd = [dev(x) for x in on], [dev(x) for x in tg]
  ad, Rd, Dd, pd = dev(a), dev(R), dev(D), dev(prob)
  # unfused reference
  qs = []
  for h, hp in ((hd[0], ond), (hd[2], tgd), (hd[1], ond)):
    q, val, adv = f(B, A), f(B, 1), f(B, A)
    _capi.call('b200rl_duelling_head_fwd', B, A, H, h.data_ptr(), 2 * H, hp[0].data_ptr(), hp[1].data_ptr(), hp[2].data_ptr(),
               hp[3].data_ptr(), val.data_ptr(), adv.data_ptr(), q.data_ptr(), st)
    qs.append(q)
  td, lps, w, pr, dq, lm = f(B), f(B), f(B), f(B), f(B, A), f(1)
  _capi.call('b200rl_dqn_td', B, A, qs[0].data_ptr(), qs[1].data_ptr(), qs[2].data_ptr(), ad.data_ptr(), Rd.data_ptr(), Dd.data_ptr(),
             pd.data_ptr(), 0.99, 1.0, 0.2, 1.0, None, 1.0 / B, td.data_ptr(), lps.data_ptr(), w.data_ptr(), pr.data_ptr(),
             dq.data_ptr(), lm.data_ptr(), 0, st)
  dval, dadv, dh = f(B, 1), f(B, A), f(B, 2 * H)
  gw = [f(H), f(1), f(A, H), f(A)]
  wsbuf = torch.empty(8 << 20, dtype=torch.uint8, device='cuda')
  _capi.call('b200rl_duelling_head_bwd', B, A, H, dq.data_ptr(), hd[0].data_ptr(), 2 * H, ond[0].data_ptr(), ond[2].data_ptr(),
             dval.data_ptr(), dadv.data_ptr(), dh.data_ptr(), 2 * H, gw[0].data_ptr(), gw[1].data_ptr(), gw[2].data_ptr(),
             gw[3].data_ptr(), wsbuf.data_ptr(), wsbuf.numel(), st)
  # fused
  wmax = torch.zeros(1, dtype=torch.float64, device='cuda')
  _capi.call('b200rl_is_weight_max', B, pd.data_ptr(), 0.2, wmax.data_ptr(), 0, st)
  q3 = f(3, B, A)
  td2, lps2, w2, pr2, dq2, dval2, dadv2, dh2 = f(B), f(B), f(B), f(B), f(B, A), f(B, 1), f(B, A), f(B, 2 * H)
  _capi.call('b200rl_dqn_head_td', B, A, H, hd[0].data_ptr(), hd[1].data_ptr(), hd[2].data_ptr(), 2 * H,
             ond[0].data_ptr(), ond[1].data_ptr(), ond[2].data_ptr(), ond[3].data_ptr(),
             tgd[0].data_ptr(), tgd[1].data_ptr(), tgd[2].data_ptr(), tgd[3].data_ptr(),
             ad.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), pd.data_ptr(), 0.99, 1.0, 0.2, 1.0, wmax.data_ptr(), 1.0 / B, 0,
             q3[0].data_ptr(), q3[1].data_ptr(), q3[2].data_ptr(), td2.data_ptr(), lps2.data_ptr(), w2.data_ptr(), pr2.data_ptr(),
             dq2.data_ptr(), dval2.data_ptr(), dadv2.data_ptr(), dh2.data_ptr(), 2 * H, 0, st)
  torch.cuda.synchronize()
  for i in range(3):
    assert torch.equal(q3[i], qs[i]), f'q[{i}]'
  for name, x, y in (('td', td2, td), ('loss', lps2, lps), ('weight', w2, w), ('priority', pr2, pr), ('dq', dq2, dq),
                     ('dval', dval2, dval), ('dadv', dadv2, dadv), ('dh', dh2, dh)):
    assert torch.equal(x, y), name
  # head parameter gradients alone
  gw2 = [f(H), f(1), f(A, H), f(A)]
  _capi.call('b200rl_duelling_head_wgrad', B, A, H, dval2.data_ptr(), dadv2.data_ptr(), hd[0].data_ptr(), 2 * H, gw2[0].data_ptr(),
             gw2[1].data_ptr(), gw2[2].data_ptr(), gw2[3].data_ptr(), wsbuf.data_ptr(), wsbuf.numel(), st)
  mean = f(1)
  _capi.call('b200rl_mean', B, lps2.data_ptr(), mean.data_ptr(), st)
  torch.cuda.synchronize()
  for x, y in zip(gw2, gw):
    assert torch.equal(x, y)
  np.testing.assert_allclose(mean.cpu().numpy()[0], lps.cpu().numpy().astype(np.float64).mean(), rtol=1e-6)


@pytest.mark.parametrize('hw', [84, 20])
def test_gather_rows_equals_gather_plus_conversion(hw):
  """b200rl_replay_gather_rows == b200rl_replay_gather followed by b200rl_conv2d_rows_bf16_from_u8, byte for byte."""
  import ctypes
  import torch
  import helpers
  from acme_b200 import _capi, replay
  rng = np.random.default_rng(hw)
  shape, A, n, B = (hw, hw, 4), 4, 3, 32
  spec, table, server, adder, oracle = helpers.make_pair(shape, np.uint8, A, n, 0.99, 0.6, max_size=200)
  for ep in range(5):
    helpers.feed_episode(rng, adder, oracle, int(rng.integers(5, 30)), n, shape, np.uint8, A)
  table.flush()
  g1, oh, pad = _geom(1, hw, 4, 8, 4, 32)
  g2, _, _ = _geom(2 * B, hw, 4, 8, 4, 32)
  fb = int(_capi.load().b200rl_conv2d_rows_bf16_bytes(ctypes.byref(g1)))
  assert fb > 0 and fb % 128 == 0
  ds = replay.ReplayDataset(table, B, seed=1)
  ds.sample_only(torch.as_tensor(rng.random(B, dtype=np.float32)).cuda())
  rows = torch.zeros(2 * B * fb, dtype=torch.uint8, device='cuda')
  ds.gather_only((rows.data_ptr(), rows.data_ptr() + B * fb, g1))
  torch.cuda.synchronize()
  o_both, R, D, a = ds.o_both.clone(), ds.R.clone(), ds.D.clone(), ds.a_tm1.clone()
  ds.o_both.zero_()
  ds.gather_only()
  want = torch.full((2 * B * fb,), 0xAB, dtype=torch.uint8, device='cuda')
  _capi.call('b200rl_conv2d_rows_bf16_from_u8', ds.o_both.data_ptr(), g2, want.data_ptr(), want.numel(), _capi.current_stream())
  torch.cuda.synchronize()
  assert torch.equal(ds.o_both, o_both) and torch.equal(ds.R, R) and torch.equal(ds.D, D) and torch.equal(ds.a_tm1, a)
  assert torch.equal(rows, want)
  server.stop()
