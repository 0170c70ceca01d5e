"""Q-network / critic / policy modules of the hot path, forward AND hand-written backward, as
sequences of calls into libb200rl (K6).  No autograd, no torch math: torch tensors are memory.

Mirrors (shapes, padding, activation placement, parameter set):
  DQNAtariNetwork   acme/tf/networks/atari.py:36-69 + duelling.py:27-59
  MLPQNetwork       examples/bsuite/run_dqn.py:46-49 (snt.Sequential([Flatten, MLP([50, 50, A])]))
  LayerNormMLP      acme/tf/networks/continuous.py:37-68
  D4PGCritic        multiplexers.py:58-80 + continuous.py:37-68 + distributional.py:36-67
  D4PGPolicy        continuous.py:30-68 + rescaling.py:55-74
Parameters live in ONE flat fp32 device buffer per network (so Adam, target copy and the NCCL
all-reduce are single passes); weights are stored [out][in] (conv: [Cout][kh][kw][Cin]) and are
exported in Sonnet's shapes (conv HWIO, linear [in, out]) by `variables()`.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

from acme_b200 import _capi
from acme_b200._capi import ACT_ELU, ACT_NONE, ACT_RELU, ACT_TANH, ConvGeom


def tf_same_pad(size: int, k: int, s: int) -> Tuple[int, int]:
  """TF 'SAME': out = ceil(size/s); pad_before = total // 2 (the remainder goes after)."""
  out = -(-size // s)
  total = max((out - 1) * s + k - size, 0)
  return out, total // 2


def _truncated_normal(rng, shape, std):
  x = rng.standard_normal(shape)
  bad = np.abs(x) > 2
  while bad.any():
    x[bad] = rng.standard_normal(int(bad.sum()))
    bad = np.abs(x) > 2
  return (x * std).astype(np.float32)


class ParamStore:
  """Flat fp32 parameter buffer + named views; gradients / Adam moments mirror the layout."""

  def __init__(self, device: int):
    self.device = device
    self.entries: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
    self.size = 0
    self.flat = None
    self.grad = None
    self.shadow = None

  def declare(self, name: str, shape):
    shape = tuple(int(s) for s in shape)
    self.entries[name] = (self.size, shape)
    n = int(np.prod(shape))
    self.size += -(-n // 8) * 8          # every tensor 32-byte aligned in fp32 and 16-byte aligned in the bf16 shadow

  def allocate(self):
    import torch
    dev = torch.device('cuda', self.device)
    self.flat = torch.zeros(self.size, dtype=torch.float32, device=dev)
    self.grad = torch.zeros(self.size, dtype=torch.float32, device=dev)

  def ensure_shadow(self):
    """bf16 copy of the parameters at the same element offsets (the weight operands of the bf16 dataflow).  Kept in
    step by the optimizer (`b200rl_adam(..., bf16_shadow)`) or `refresh_shadow()` after any other parameter write."""
    import torch
    if self.shadow is None:
      self.shadow = torch.zeros(self.size, dtype=torch.bfloat16, device=self.flat.device)
      self.refresh_shadow()
    return self.shadow

  def refresh_shadow(self, off: int = 0, n: Optional[int] = None):
    if self.shadow is None:
      return
    n = self.size - off if n is None else n
    assert off % 8 == 0 and n % 8 == 0
    _capi.call('b200rl_bf16_from_f32', n, self.flat.data_ptr() + 4 * off, self.shadow.data_ptr() + 2 * off,
               _capi.current_stream())

  def sp(self, name: str) -> int:
    """Device address of `name` in the bf16 shadow."""
    return self.shadow.data_ptr() + 2 * self.entries[name][0]

  def rebind(self, flat, grad):
    """Moves the parameters / gradients into caller-provided storage (the peer-mapped region of
    `parallel.PeerExchange`); call before any CUDA graph that uses them is captured."""
    assert flat.numel() == self.size and grad.numel() == self.size
    flat.copy_(self.flat)
    grad.copy_(self.grad)
    self.flat, self.grad = flat, grad

  def view(self, name: str, grad: bool = False):
    off, shape = self.entries[name]
    n = int(np.prod(shape))
    return (self.grad if grad else self.flat)[off:off + n].view(shape)

  def p(self, name: str) -> int:
    return self.flat.data_ptr() + 4 * self.entries[name][0]

  def g(self, name: str) -> int:
    return self.grad.data_ptr() + 4 * self.entries[name][0]

  def set(self, name: str, value: np.ndarray):
    import torch
    self.view(name).copy_(torch.as_tensor(np.ascontiguousarray(value, dtype=np.float32)))
    if self.shadow is not None:
      self.refresh_shadow()

  def get(self, name: str, grad: bool = False) -> np.ndarray:
    return self.view(name, grad).detach().cpu().numpy().copy()


class Marks:
  """tools/step_phases.py: global-timer stamps issued between kernels (also inside a captured step); name -> slot."""

  def __init__(self, device: int, slots: int = 128):
    import torch
    self.buf = torch.zeros(slots, dtype=torch.int64, device=torch.device('cuda', device))
    self.slots = {}

  def mark(self, name: str):
    import ctypes
    slot = self.slots.setdefault(name, len(self.slots))
    if slot < self.buf.numel():
      _capi.load().b200rl_debug_stamp(ctypes.c_void_p(self.buf.data_ptr()), slot, ctypes.c_void_p(_capi.current_stream()))

  def timeline(self):
    t = self.buf.cpu().numpy()
    ev = sorted((int(t[s]), n) for n, s in self.slots.items() if s < len(t) and t[s] > 0)
    t0 = ev[0][0] if ev else 0
    return [(n, (x - t0) / 1e3) for x, n in ev]


class Network:
  """Common plumbing: parameter store, workspace, precision."""
  marks: Optional[Marks] = None
  mark_prefix = ''

  def _mark(self, name: str):
    if self.marks is not None:
      self.marks.mark(self.mark_prefix + name)

  def __init__(self, device: int = 0, precision: int = _capi.PRECISION_FP32):
    self.device = device
    self.precision = precision
    self.params = ParamStore(device)
    self._ws = None

  def _finalize(self):
    import torch
    _capi.require_device(self.device)
    self.params.allocate()
    # one split-K / partial-sum workspace per concurrent stream (see Network.lanes)
    self._ws_all = [torch.empty(96 << 20, dtype=torch.uint8, device=torch.device('cuda', self.device)) for _ in range(3)]
    self._lane = 0

  @property
  def ws(self):
    w = self._ws_all[self._lane]
    return w.data_ptr(), w.numel()

  def lane(self, i: int):
    """Selects the workspace used by the following calls; kernels issued on different streams must use
    different lanes (0..2)."""
    self._lane = i
    return self

  def copy_params_from(self, other: 'Network'):
    self.params.flat.copy_(other.params.flat)
    self.params.refresh_shadow()

  def clone(self) -> 'Network':
    """Same architecture, separately stored copy of the parameters (copy.deepcopy(network),
    `acme/agents/tf/dqn/agent.py:127`).  The split-K workspace is shared (same stream)."""
    import copy
    other = copy.copy(self)
    other.params = ParamStore(self.device)
    other.params.entries = dict(self.params.entries)
    other.params.size = self.params.size
    other.params.allocate()
    other.params.flat.copy_(self.params.flat)
    if self.params.shadow is not None:
      other.params.ensure_shadow()
    if hasattr(self, '_rows'):
      other._rows = {}                  # row images are per network (their forwards may run concurrently)
    return other

  # --- Sonnet-shaped export / import (get_variables, checkpoints, parity tests)
  def variables(self, grad: bool = False) -> Dict[str, np.ndarray]:
    out = {}
    for sonnet_name, (name, kind) in self._export.items():
      a = self.params.get(name, grad)
      if kind == 'conv':      # OHWI -> HWIO
        a = a.transpose(1, 2, 3, 0)
      elif kind == 'linear':  # [out, in] -> [in, out]
        a = a.T
      elif isinstance(kind, tuple) and kind[0] == 'rows':   # rows [lo:hi] of a stacked [out, in]
        a = a[kind[1]:kind[2]].T if a.ndim == 2 else a[kind[1]:kind[2]]
      out[sonnet_name] = np.ascontiguousarray(a)
    return out

  def load_variables(self, variables: Dict[str, np.ndarray]):
    staged: Dict[str, np.ndarray] = {}
    for sonnet_name, (name, kind) in self._export.items():
      v = np.asarray(variables[sonnet_name], np.float32)
      if kind == 'conv':
        staged[name] = v.transpose(3, 0, 1, 2)
      elif kind == 'linear':
        staged[name] = v.T
      elif isinstance(kind, tuple) and kind[0] == 'rows':
        full = staged.get(name)
        if full is None:
          full = self.params.get(name)
        full[kind[1]:kind[2]] = v.T if full.ndim == 2 else v
        staged[name] = full
      else:
        staged[name] = v
    for name, v in staged.items():
      self.params.set(name, v)


def _linear(M, N, K, x, ldx, w, b, y, ldy, act, net: Network):
  ws, wsb = net.ws
  _capi.call('b200rl_linear_fwd', M, N, K, x, ldx, w, b, y, ldy, act, net.precision, ws, wsb,
             _capi.current_stream())


def _linear_dgrad(M, N, K, dy, lddy, w, dx, lddx, mask, mask_act, net: Network):
  ws, wsb = net.ws
  _capi.call('b200rl_linear_dgrad', M, N, K, dy, lddy, w, dx, lddx, mask, mask_act, net.precision, ws,
             wsb, _capi.current_stream())


def _linear_wgrad(M, N, K, dy, lddy, x, ldx, dw, db, net: Network):
  ws, wsb = net.ws
  _capi.call('b200rl_linear_wgrad', M, N, K, dy, lddy, x, ldx, dw, db, net.precision, ws, wsb,
             _capi.current_stream())


# =============================================================================== DQN Atari network
class DQNAtariNetwork(Network):
  """AtariTorso (3 SAME-padded convs + ReLU, NHWC) -> DuellingMLP([512]) (atari.py:36-69)."""

  def __init__(self, num_actions: int, device: int = 0, precision: int = _capi.PRECISION_FP32,
               seed: int = 0, input_hw: int = 84, input_channels: int = 4):
    super().__init__(device, precision)
    self.A = int(num_actions)
    self.in_hw, self.in_c = input_hw, input_channels
    self.convs = []  # (k, stride, cin, cout, in_hw, out_hw, pad)
    h, c = input_hw, input_channels
    for k, s, co in ((8, 4, 32), (4, 2, 64), (3, 1, 64)):
      oh, pad = tf_same_pad(h, k, s)
      self.convs.append((k, s, c, co, h, oh, pad))
      h, c = oh, co
    self.flat_dim = h * h * c
    P = self.params
    for i, (k, s, ci, co, _, _, _) in enumerate(self.convs, 1):
      P.declare(f'conv{i}.w', (co, k, k, ci))
      P.declare(f'conv{i}.b', (co,))
    P.declare('fc1.w', (1024, self.flat_dim))   # rows 0..511 value stream, 512..1023 advantage stream
    P.declare('fc1.b', (1024,))
    P.declare('v2.w', (1, 512))
    P.declare('v2.b', (1,))
    P.declare('a2.w', (self.A, 512))
    P.declare('a2.b', (self.A,))
    self._export = {
        'conv1/w': ('conv1.w', 'conv'), 'conv1/b': ('conv1.b', None),
        'conv2/w': ('conv2.w', 'conv'), 'conv2/b': ('conv2.b', None),
        'conv3/w': ('conv3.w', 'conv'), 'conv3/b': ('conv3.b', None),
        'value/l0/w': ('fc1.w', ('rows', 0, 512)), 'value/l0/b': ('fc1.b', ('rows', 0, 512)),
        'value/l1/w': ('v2.w', 'linear'), 'value/l1/b': ('v2.b', None),
        'adv/l0/w': ('fc1.w', ('rows', 512, 1024)), 'adv/l0/b': ('fc1.b', ('rows', 512, 1024)),
        'adv/l1/w': ('a2.w', 'linear'), 'adv/l1/b': ('a2.b', None),
    }
    self._finalize()
    self.init(seed)

  def init(self, seed: int):
    """Sonnet defaults: TruncatedNormal(stddev=1/sqrt(fan_in)) weights, zero biases."""
    rng = np.random.default_rng(seed)
    P = self.params
    for i, (k, s, ci, co, _, _, _) in enumerate(self.convs, 1):
      P.set(f'conv{i}.w', _truncated_normal(rng, (co, k, k, ci), 1 / math.sqrt(k * k * ci)))
    P.set('fc1.w', _truncated_normal(rng, (1024, self.flat_dim), 1 / math.sqrt(self.flat_dim)))
    P.set('v2.w', _truncated_normal(rng, (1, 512), 1 / math.sqrt(512)))
    P.set('a2.w', _truncated_normal(rng, (self.A, 512), 1 / math.sqrt(512)))

  def geom(self, i: int, B: int) -> ConvGeom:
    k, s, ci, co, h, oh, pad = self.convs[i]
    return ConvGeom(B=B, H=h, W=h, C=ci, kh=k, kw=k, stride=s, pad_top=pad, pad_left=pad, OH=oh, OW=oh, Cout=co)

  # ---- buffers.  In the bf16 dataflow (precision 2) activations y1..y3 and the back-propagated gradients dh, dy1..dy3
  # are bf16; the hidden row h (consumed by the duelling head in fp32), q, value / advantage and everything the loss
  # touches stay fp32.
  @property
  def flow(self) -> bool:
    return self.precision == _capi.PRECISION_BF16FLOW

  def make_buffers(self, B: int):
    import torch
    dev = torch.device('cuda', self.device)
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    a = (lambda *shape: torch.empty(shape, dtype=torch.bfloat16, device=dev)) if self.flow else f
    bufs = dict(B=B, q=f(B, self.A), h=f(B, 1024), val=f(B, 1), adv=f(B, self.A))
    for i, (k, s, ci, co, h, oh, pad) in enumerate(self.convs, 1):
      bufs[f'y{i}'] = a(B, oh, oh, co)
    return bufs

  def make_grad_buffers(self, B: int):
    import torch
    dev = torch.device('cuda', self.device)
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    a = (lambda *shape: torch.empty(shape, dtype=torch.bfloat16, device=dev)) if self.flow else f
    g = dict(dh=a(B, 1024), dval=f(B, 1), dadv=f(B, self.A))
    if self.flow:
      g['dh32'] = f(B, 1024)           # the unfused head backward writes fp32; converted for the bf16 GEMMs
    for i, (k, s, ci, co, h, oh, pad) in enumerate(self.convs, 1):
      g[f'dy{i}'] = a(B, oh, oh, co)
    return g

  def rows_frame_bytes(self) -> int:
    """Bytes of ONE frame in the bf16 row image (bf16 dataflow): a batched image can be sliced per sample range."""
    import ctypes
    return int(_capi.load().b200rl_conv2d_rows_bf16_bytes(ctypes.byref(self.geom(0, 1)))) if self.flow else 0

  def rows_buffer(self, B: int, slot: str = 'x'):
    """The (cached, zero-initialised) row-image buffer of B frames; None if the geometry has no row image.  Zeroed because
    `b200rl_replay_gather_rows` writes only the frames' interior and relies on the padding staying zero."""
    import ctypes
    import torch
    cache = self.__dict__.setdefault('_rows', {})
    if (slot, B) not in cache:
      lib = _capi.load()
      fn = lib.b200rl_conv2d_rows_bf16_bytes if self.flow else lib.b200rl_conv2d_rows_bytes
      nbytes = int(fn(ctypes.byref(self.geom(0, B))))
      cache[(slot, B)] = (torch.zeros(nbytes, dtype=torch.uint8, device=torch.device('cuda', self.device)) if nbytes > 0 else None)
    return cache[(slot, B)]

  def prepare_frames(self, obs, slot: str = 'x'):
    """uint8 frames -> the row image that first-layer calls accept, built ONCE for all the passes over the same frames
    (two networks on o_t; forward + weight gradient on o_tm1).  Precision 1: zero-padded fp32 rows x/255
    (`b200rl_conv2d_rows_from_u8`); precision 2: zero-padded bf16 rows of the integer pixel values
    (`b200rl_conv2d_rows_bf16_from_u8`).  None when not applicable (fp32 mode, float frames, geometry not eligible):
    callers then pass the frames themselves."""
    import ctypes
    import torch
    if self.precision == _capi.PRECISION_FP32 or obs.dtype != torch.uint8:
      return None
    B = obs.shape[0]
    buf = self.rows_buffer(B, slot)
    if buf is None:
      if self.flow:
        raise ValueError('the bf16 dataflow needs the Atari first-layer geometry (C = 4, kw * C = 32)')
      return None
    name = 'b200rl_conv2d_rows_bf16_from_u8' if self.flow else 'b200rl_conv2d_rows_from_u8'
    _capi.call(name, obs.data_ptr(), self.geom(0, B), buf.data_ptr(), buf.numel(), _capi.current_stream())
    return buf

  def features(self, obs, bufs, before_fc1=None, rows=None):
    """The torso and the first dense layer of both streams: obs -> bufs['h'] (post-ReLU, fp32 [B, 1024]).
    obs: uint8 or float32 [B, H, W, C] (NHWC); uint8 is read as float32(x)/255.  `rows`: the frames' row image from
    prepare_frames() (required in the bf16 dataflow; built here if absent).  `before_fc1` (optional callable) runs after
    the torso has been issued and before the dense layer: the pipelined data-parallel learner makes the stream wait there
    for the fc1 + head parameters, which arrive while the convolutions run."""
    import torch
    B, P, st = bufs['B'], self.params, _capi.current_stream()
    ws, wsb = self.ws
    if self.flow:
      P.ensure_shadow()
      if rows is None:
        if obs.dtype != torch.uint8:
          raise ValueError('the bf16 dataflow reads uint8 frames')
        rows = self.prepare_frames(obs, 'fwd')
      x, x_rows = rows.data_ptr(), 1
      self._mark(f'fwd{B}.start')
      for i in range(3):
        y = bufs[f'y{i + 1}']
        _capi.call('b200rl_conv2d_fwd_bf16', x, x_rows, P.sp(f'conv{i + 1}.w'), P.p(f'conv{i + 1}.b'), y.data_ptr(), 1,
                   self.geom(i, B), ACT_RELU, ws, wsb, st)
        self._mark(f'fwd{B}.conv{i + 1}')
        x, x_rows = y.data_ptr(), 0
      if before_fc1 is not None:
        before_fc1()
      _capi.call('b200rl_linear_fwd_bf16', B, 1024, self.flat_dim, x, self.flat_dim, P.sp('fc1.w'), P.p('fc1.b'),
                 bufs['h'].data_ptr(), 1024, 0, ACT_RELU, ws, wsb, st)
      self._mark(f'fwd{B}.fc1')
      return bufs['h']
    x, x_u8 = obs.data_ptr(), int(obs.dtype == torch.uint8)
    if rows is not None:              # the frames' row image from prepare_frames()
      x, x_u8 = rows.data_ptr(), 2
    for i in range(3):
      y = bufs[f'y{i + 1}']
      g = self.geom(i, B)
      _capi.call('b200rl_conv2d_fwd', x, x_u8, P.p(f'conv{i + 1}.w'), P.p(f'conv{i + 1}.b'), y.data_ptr(),
                 g, ACT_RELU, self.precision, ws, wsb, st)
      x, x_u8 = y.data_ptr(), 0
    if before_fc1 is not None:
      before_fc1()
    h = bufs['h']
    _linear(B, 1024, self.flat_dim, x, self.flat_dim, P.p('fc1.w'), P.p('fc1.b'), h.data_ptr(), 1024, ACT_RELU, self)
    return h

  def head_params(self):
    """(wv, bv, wa, ba) device addresses of the duelling head (fp32 master weights)."""
    P = self.params
    return P.p('v2.w'), P.p('v2.b'), P.p('a2.w'), P.p('a2.b')

  def forward(self, obs, bufs, before_fc1=None, rows=None) -> 'torch.Tensor':
    """features() followed by the duelling head (`duelling.py:37-59`): q [B, A]."""
    P, B = self.params, bufs['B']
    h = self.features(obs, bufs, before_fc1=before_fc1, rows=rows)
    _capi.call('b200rl_duelling_head_fwd', B, self.A, 512, h.data_ptr(), 1024, P.p('v2.w'), P.p('v2.b'), P.p('a2.w'),
               P.p('a2.b'), bufs['val'].data_ptr(), bufs['adv'].data_ptr(), bufs['q'].data_ptr(), _capi.current_stream())
    return bufs['q']

  def _head_backward(self, bufs, gbufs, dq):
    """Unfused duelling-head backward: dval, dadv, dh (masked by relu'(h)) and the head's four parameter gradients."""
    B, P = bufs['B'], self.params
    ws, wsb = self.ws
    dh = gbufs['dh32'] if self.flow else gbufs['dh']
    _capi.call('b200rl_duelling_head_bwd', B, self.A, 512, dq.data_ptr(), bufs['h'].data_ptr(), 1024, P.p('v2.w'), P.p('a2.w'),
               gbufs['dval'].data_ptr(), gbufs['dadv'].data_ptr(), dh.data_ptr(), 1024, P.g('v2.w'), P.g('v2.b'), P.g('a2.w'),
               P.g('a2.b'), ws, wsb, _capi.current_stream())
    if self.flow:
      _capi.call('b200rl_bf16_from_f32', B * 1024, dh.data_ptr(), gbufs['dh'].data_ptr(), _capi.current_stream())

  def head_wgrad(self, bufs, gbufs):
    """The head's parameter gradients from gbufs['dval'] / ['dadv'] (after the fused head + TD kernel)."""
    B, P = bufs['B'], self.params
    ws, wsb = self.ws
    _capi.call('b200rl_duelling_head_wgrad', B, self.A, 512, gbufs['dval'].data_ptr(), gbufs['dadv'].data_ptr(),
               bufs['h'].data_ptr(), 1024, P.g('v2.w'), P.g('v2.b'), P.g('a2.w'), P.g('a2.b'), ws, wsb, _capi.current_stream())

  # layer calls of the backward pass in either dataflow
  def _fc1_bias_grad(self, bufs, gbufs):
    """bf16 dataflow: d fc1.b = column sums of dh, as its own launch."""
    B, P = bufs['B'], self.params
    ws, wsb = self.ws
    _capi.call('b200rl_colsum_bf16', B, 1024, gbufs['dh'].data_ptr(), 1024, P.g('fc1.b'), ws, wsb, _capi.current_stream())

  def _fc1_wgrad(self, bufs, gbufs, bias: bool = True):
    B, P = bufs['B'], self.params
    ws, wsb = self.ws
    if self.flow:
      _capi.call('b200rl_linear_wgrad_bf16', B, 1024, self.flat_dim, gbufs['dh'].data_ptr(), 1024, bufs['y3'].data_ptr(),
                 self.flat_dim, P.g('fc1.w'), P.g('fc1.b') if bias else None, ws, wsb, _capi.current_stream())
    else:
      _linear_wgrad(B, 1024, self.flat_dim, gbufs['dh'].data_ptr(), 1024, bufs['y3'].data_ptr(), self.flat_dim, P.g('fc1.w'),
                    P.g('fc1.b'), self)

  def _fc1_dgrad(self, bufs, gbufs):
    B, P = bufs['B'], self.params
    ws, wsb = self.ws
    if self.flow:
      _capi.call('b200rl_linear_dgrad_bf16', B, 1024, self.flat_dim, gbufs['dh'].data_ptr(), 1024, P.sp('fc1.w'),
                 gbufs['dy3'].data_ptr(), self.flat_dim, 1, bufs['y3'].data_ptr(), 1, ACT_RELU, ws, wsb, _capi.current_stream())
    else:
      _linear_dgrad(B, 1024, self.flat_dim, gbufs['dh'].data_ptr(), 1024, P.p('fc1.w'), gbufs['dy3'].data_ptr(), self.flat_dim,
                    bufs['y3'].data_ptr(), ACT_RELU, self)

  def _conv_bias_grad(self, i, bufs, gbufs):
    """bf16 dataflow: db of conv layer i+1 = column sums of dy, as its own launch (so it can sit on another stream)."""
    B, P = bufs['B'], self.params
    ws, wsb = self.ws
    _, _, _, co, _, oh, _ = self.convs[i]
    _capi.call('b200rl_colsum_bf16', B * oh * oh, co, gbufs[f'dy{i + 1}'].data_ptr(), co, P.g(f'conv{i + 1}.b'), ws, wsb,
               _capi.current_stream())

  def _conv_wgrad(self, i, obs, bufs, gbufs, rows, bias: bool = True):
    import torch
    B, P = bufs['B'], self.params
    ws, wsb = self.ws
    g = self.geom(i, B)
    dy = gbufs[f'dy{i + 1}'].data_ptr()
    if self.flow:
      if i > 0:
        x, x_rows = bufs[f'y{i}'].data_ptr(), 0
      else:
        if rows is None:
          rows = self.prepare_frames(obs, 'bwd')
        x, x_rows = rows.data_ptr(), 1
      _capi.call('b200rl_conv2d_wgrad_bf16', x, x_rows, dy, P.g(f'conv{i + 1}.w'), P.g(f'conv{i + 1}.b') if bias else None, g,
                 ws, wsb, _capi.current_stream())
      return
    if i > 0:
      x, x_u8 = bufs[f'y{i}'].data_ptr(), 0
    elif rows is not None:
      x, x_u8 = rows.data_ptr(), 2
    else:
      x, x_u8 = obs.data_ptr(), int(obs.dtype == torch.uint8)
    _capi.call('b200rl_conv2d_wgrad', x, x_u8, dy, P.g(f'conv{i + 1}.w'), P.g(f'conv{i + 1}.b'), g, self.precision, ws, wsb,
               _capi.current_stream())

  def _conv_dgrad(self, i, bufs, gbufs):
    B, P = bufs['B'], self.params
    ws, wsb = self.ws
    g = self.geom(i, B)
    dy = gbufs[f'dy{i + 1}'].data_ptr()
    if self.flow:
      _capi.call('b200rl_conv2d_dgrad_bf16', dy, P.sp(f'conv{i + 1}.w'), gbufs[f'dy{i}'].data_ptr(), 1, g,
                 bufs[f'y{i}'].data_ptr(), 1, ACT_RELU, _capi.current_stream())
    else:
      _capi.call('b200rl_conv2d_dgrad', dy, P.p(f'conv{i + 1}.w'), gbufs[f'dy{i}'].data_ptr(), g, bufs[f'y{i}'].data_ptr(),
                 ACT_RELU, self.precision, ws, wsb, _capi.current_stream())

  def backward(self, obs, bufs, gbufs, dq, side_stream=None, rows=None):
    """Accumulates nothing: overwrites params.grad with d(loss)/d(params) given dq [B, A].  `rows`: the row image of
    `obs` from prepare_frames(), if the caller has one.

    With `side_stream`, the weight-gradient GEMMs of fc1 / conv3 / conv2 run on it (workspace lane 1)
    concurrently with the data-gradient chain on the current stream: the two are independent once a
    layer's dy exists, and each kernel alone is too latency-bound to fill the machine."""
    if side_stream is not None:
      return self._backward_two_streams(obs, bufs, gbufs, dq, side_stream, rows)
    self._head_backward(bufs, gbufs, dq)
    self._fc1_wgrad(bufs, gbufs)
    self._fc1_dgrad(bufs, gbufs)
    for i in (2, 1, 0):
      self._conv_wgrad(i, obs, bufs, gbufs, rows)
      if i > 0:
        self._conv_dgrad(i, bufs, gbufs)

  def _backward_two_streams(self, obs, bufs, gbufs, dq, side, rows=None):
    self.backward_dense_part(bufs, gbufs, dq, side)
    self.backward_conv_part(obs, bufs, gbufs, side, rows)

  def backward_after_head(self, obs, bufs, gbufs, s1, s2, rows=None, on_dense_done=None):
    """Everything behind the fused head + TD kernel (which has produced gbufs['dh'] / ['dval'] / ['dadv']), scheduled on
    three streams: the data-gradient chain fc1 -> conv3 -> conv2 (+ conv1's weight gradient) on the current stream, the
    weight-gradient GEMMs on `s1`, the head's parameter gradients and (bf16 dataflow) the bias-gradient column sums on
    `s2`.  Workspace lanes 0 / 1 / 2.  `on_dense_done(event_list)` is called once fc1's and the head's gradients and fc1's
    data gradient (the last reader of fc1's weights) have been issued, with the events that mark their completion: a
    caller may start the optimizer update of that bucket there, underneath the convolution backward."""
    import torch
    main = torch.cuda.current_stream()
    flow = self.flow

    def fork(*streams):
      ev = torch.cuda.Event()
      ev.record(main)
      for st in streams:
        st.wait_event(ev)

    self._mark('bwd.start')
    fork(s1, s2)
    with torch.cuda.stream(s1):
      self.lane(1)
      self._fc1_wgrad(bufs, gbufs, bias=not flow)
      self._mark('bwd.s1.fc1_wgrad')
      ev_w = torch.cuda.Event()
      ev_w.record(s1)
    with torch.cuda.stream(s2):
      self.lane(2)
      self.head_wgrad(bufs, gbufs)
      if flow:
        self._fc1_bias_grad(bufs, gbufs)
      self._mark('bwd.s2.head_wgrad+fc1_db')
      ev_h = torch.cuda.Event()
      ev_h.record(s2)
    self.lane(0)
    self._fc1_dgrad(bufs, gbufs)
    self._mark('bwd.fc1_dgrad')
    if on_dense_done is not None:
      # fc1's data gradient READS fc1's weights: an optimizer update of that bucket must wait for it too
      ev_d = torch.cuda.Event()
      ev_d.record(main)
      on_dense_done([ev_w, ev_h, ev_d])
    for i in (2, 1):
      fork(s1, s2) if flow else fork(s1)
      with torch.cuda.stream(s1):
        self.lane(1)
        self._conv_wgrad(i, obs, bufs, gbufs, rows, bias=not flow)
        self._mark(f'bwd.s1.conv{i + 1}_wgrad')
      if flow:
        with torch.cuda.stream(s2):
          self.lane(2)
          self._conv_bias_grad(i, bufs, gbufs)
          self._mark(f'bwd.s2.conv{i + 1}_db')
      self.lane(0)
      self._conv_dgrad(i, bufs, gbufs)
      self._mark(f'bwd.conv{i + 1}_dgrad')
    if flow:
      fork(s2)
      with torch.cuda.stream(s2):
        self.lane(2)
        self._conv_bias_grad(0, bufs, gbufs)
    self.lane(0)
    self._conv_wgrad(0, obs, bufs, gbufs, rows, bias=not flow)
    self._mark('bwd.conv1_wgrad')
    for st in (s1, s2):
      ev = torch.cuda.Event()
      ev.record(st)
      main.wait_event(ev)

  def grad_buckets(self):
    """(offset, count) in floats of the gradient regions that become final after backward_dense_part
    (fc1 + heads: the tail of the flat buffer, 99% of the bytes) and after backward_conv_part (the convs)."""
    split = self.params.entries['fc1.w'][0]
    return (split, self.params.size - split), (0, split)

  def backward_dense_part(self, bufs, gbufs, dq, side):
    """Duelling head + fc1: parameter gradients of everything after the torso, and dy3.  dq = None: the fused head + TD
    kernel has already produced gbufs['dh'] / ['dval'] / ['dadv']; only the head's parameter gradients remain (side)."""
    import torch
    main = torch.cuda.current_stream()
    self.lane(0)
    if dq is not None:
      self._head_backward(bufs, gbufs, dq)
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    with torch.cuda.stream(side):
      self.lane(1)
      if dq is None:
        self.head_wgrad(bufs, gbufs)
      self._fc1_wgrad(bufs, gbufs)
    self.lane(0)
    self._fc1_dgrad(bufs, gbufs)
    ev = torch.cuda.Event()
    ev.record(side)
    main.wait_event(ev)

  def backward_conv_part(self, obs, bufs, gbufs, side, rows=None):
    """The three convolutions: weight gradients on `side`, data gradients on the current stream."""
    import torch
    main = torch.cuda.current_stream()
    for i in (2, 1, 0):
      if i > 0:
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
          self.lane(1)
          self._conv_wgrad(i, obs, bufs, gbufs, rows)
        self.lane(0)
        self._conv_dgrad(i, bufs, gbufs)
      else:   # conv1 has no data gradient: its weight gradient finishes the main chain
        if self.flow:   # ... and its bias gradient runs beside it
          ev = torch.cuda.Event()
          ev.record(main)
          side.wait_event(ev)
          with torch.cuda.stream(side):
            self.lane(1)
            self._conv_bias_grad(0, bufs, gbufs)
          self.lane(0)
        self._conv_wgrad(0, obs, bufs, gbufs, rows, bias=not self.flow)
    ev = torch.cuda.Event()
    ev.record(side)
    main.wait_event(ev)
    self.lane(0)


# =============================================================================== plain MLP Q-net
class MLPQNetwork(Network):
  """snt.Sequential([snt.Flatten(), snt.nets.MLP(sizes)]) with ReLU between layers."""

  def __init__(self, input_dim: int, sizes, device: int = 0, precision: int = _capi.PRECISION_FP32, seed: int = 0):
    super().__init__(device, precision)
    self.in_dim = int(input_dim)
    self.sizes = [int(s) for s in sizes]
    self.A = self.sizes[-1]
    d = self.in_dim
    self._export = {}
    for i, hdim in enumerate(self.sizes):
      self.params.declare(f'l{i}.w', (hdim, d))
      self.params.declare(f'l{i}.b', (hdim,))
      self._export[f'l{i}/w'] = (f'l{i}.w', 'linear')
      self._export[f'l{i}/b'] = (f'l{i}.b', None)
      d = hdim
    self._finalize()
    rng = np.random.default_rng(seed)
    d = self.in_dim
    for i, hdim in enumerate(self.sizes):
      self.params.set(f'l{i}.w', _truncated_normal(rng, (hdim, d), 1 / math.sqrt(d)))
      d = hdim

  def make_buffers(self, B: int):
    import torch
    dev = torch.device('cuda', self.device)
    bufs = dict(B=B)
    for i, hdim in enumerate(self.sizes):
      bufs[f'y{i}'] = torch.empty((B, hdim), dtype=torch.float32, device=dev)
    bufs['q'] = bufs[f'y{len(self.sizes) - 1}']
    return bufs

  def make_grad_buffers(self, B: int):
    import torch
    dev = torch.device('cuda', self.device)
    return {f'dy{i}': torch.empty((B, hdim), dtype=torch.float32, device=dev) for i, hdim in enumerate(self.sizes)}

  def forward(self, obs, bufs):
    import torch
    if obs.dtype != torch.float32:
      raise ValueError('MLPQNetwork expects float32 observations')
    B, P = bufs['B'], self.params
    x, d = obs.data_ptr(), self.in_dim
    n = len(self.sizes)
    for i, hdim in enumerate(self.sizes):
      y = bufs[f'y{i}']
      _linear(B, hdim, d, x, d, P.p(f'l{i}.w'), P.p(f'l{i}.b'), y.data_ptr(), hdim,
              ACT_RELU if i + 1 < n else ACT_NONE, self)
      x, d = y.data_ptr(), hdim
    return bufs['q']

  def backward(self, obs, bufs, gbufs, dq):
    B, P = bufs['B'], self.params
    n = len(self.sizes)
    dy = dq.data_ptr()
    for i in range(n - 1, -1, -1):
      hdim = self.sizes[i]
      d = self.sizes[i - 1] if i > 0 else self.in_dim
      x = bufs[f'y{i - 1}'].data_ptr() if i > 0 else obs.data_ptr()
      _linear_wgrad(B, hdim, d, dy, hdim, x, d, P.g(f'l{i}.w'), P.g(f'l{i}.b'), self)
      if i > 0:
        dx = gbufs[f'dy{i - 1}'].data_ptr()
        _linear_dgrad(B, hdim, d, dy, hdim, P.p(f'l{i}.w'), dx, d, x, ACT_RELU, self)
        dy = dx


# =============================================================================== D4PG networks
class LayerNormMLP(Network):
  """Linear -> LayerNorm -> tanh -> [Linear -> ELU]* -> optional linear head (continuous.py:37-68)."""

  def __init__(self, input_dim: int, sizes, head_dim: int, head_name: str, head_scale: Optional[float],
               device: int, precision: int, seed: int):
    super().__init__(device, precision)
    self.in_dim, self.sizes, self.head_dim, self.head = int(input_dim), [int(s) for s in sizes], int(head_dim), head_name
    P = self.params
    d = self.in_dim
    self._export = {}
    for i, hdim in enumerate(self.sizes):
      P.declare(f'l{i}.w', (hdim, d))
      P.declare(f'l{i}.b', (hdim,))
      self._export[f'l{i}/w'] = (f'l{i}.w', 'linear')
      self._export[f'l{i}/b'] = (f'l{i}.b', None)
      if i == 0:
        P.declare('ln.scale', (hdim,))
        P.declare('ln.offset', (hdim,))
        self._export['ln/scale'] = ('ln.scale', None)
        self._export['ln/offset'] = ('ln.offset', None)
      d = hdim
    P.declare(f'{head_name}.w', (self.head_dim, d))
    P.declare(f'{head_name}.b', (self.head_dim,))
    self._export[f'{head_name}/w'] = (f'{head_name}.w', 'linear')
    self._export[f'{head_name}/b'] = (f'{head_name}.b', None)
    self._finalize()
    rng = np.random.default_rng(seed)
    d = self.in_dim
    for i, hdim in enumerate(self.sizes):
      lim = math.sqrt(3.0 * 0.333 / hdim)   # VarianceScaling(0.333, 'fan_out', 'uniform')
      P.set(f'l{i}.w', rng.uniform(-lim, lim, (hdim, d)).astype(np.float32))
      d = hdim
    P.set('ln.scale', np.ones(self.sizes[0], np.float32))
    if head_scale is None:
      P.set(f'{head_name}.w', _truncated_normal(rng, (self.head_dim, d), 1 / math.sqrt(d)))
    else:  # NearZeroInitializedLinear: VarianceScaling(scale) -> truncated normal over fan_in
      P.set(f'{head_name}.w', _truncated_normal(rng, (self.head_dim, d), math.sqrt(head_scale / d) / .87962566103423978))

  def make_buffers(self, B: int):
    import torch
    dev = torch.device('cuda', self.device)
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    h0 = self.sizes[0]
    bufs = dict(B=B, x=f(B, self.in_dim), z0=f(B, h0), y0=f(B, h0), xhat=f(B, h0), rstd=f(B), out=f(B, self.head_dim))
    for i in range(1, len(self.sizes)):
      bufs[f'y{i}'] = f(B, self.sizes[i])
    return bufs

  def make_grad_buffers(self, B: int):
    import torch
    dev = torch.device('cuda', self.device)
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    g = dict(dz0=f(B, self.sizes[0]), dx=f(B, self.in_dim))
    for i in range(len(self.sizes)):
      g[f'dy{i}'] = f(B, self.sizes[i])
    return g

  def forward_flat(self, x_ptr: int, bufs):
    B, P, st = bufs['B'], self.params, _capi.current_stream()
    h0 = self.sizes[0]
    _linear(B, h0, self.in_dim, x_ptr, self.in_dim, P.p('l0.w'), P.p('l0.b'), bufs['z0'].data_ptr(), h0, ACT_NONE, self)
    _capi.call('b200rl_layernorm_tanh_fwd', B, h0, bufs['z0'].data_ptr(), P.p('ln.scale'), P.p('ln.offset'), 1e-5,
               bufs['y0'].data_ptr(), bufs['xhat'].data_ptr(), bufs['rstd'].data_ptr(), st)
    x, d = bufs['y0'].data_ptr(), h0
    for i in range(1, len(self.sizes)):
      y = bufs[f'y{i}']
      _linear(B, self.sizes[i], d, x, d, P.p(f'l{i}.w'), P.p(f'l{i}.b'), y.data_ptr(), self.sizes[i], ACT_ELU, self)
      x, d = y.data_ptr(), self.sizes[i]
    _linear(B, self.head_dim, d, x, d, P.p(f'{self.head}.w'), P.p(f'{self.head}.b'), bufs['out'].data_ptr(),
            self.head_dim, ACT_NONE, self)
    return bufs['out']

  def backward_flat(self, x_ptr: int, bufs, gbufs, dout_ptr: int, param_grads: bool = True, input_grad: bool = False):
    """dout: gradient w.r.t. the head output.  Writes params.grad (if param_grads) and gbufs['dx']
    (if input_grad)."""
    B, P, st = bufs['B'], self.params, _capi.current_stream()
    n = len(self.sizes)
    d_last = self.sizes[-1]
    last = bufs[f'y{n - 1}'] if n > 1 else bufs['y0']
    if param_grads:
      _linear_wgrad(B, self.head_dim, d_last, dout_ptr, self.head_dim, last.data_ptr(), d_last,
                    P.g(f'{self.head}.w'), P.g(f'{self.head}.b'), self)
    dy = gbufs[f'dy{n - 1}'].data_ptr()
    _linear_dgrad(B, self.head_dim, d_last, dout_ptr, self.head_dim, P.p(f'{self.head}.w'), dy, d_last,
                  last.data_ptr(), ACT_ELU if n > 1 else ACT_NONE, self)
    for i in range(n - 1, 0, -1):
      hdim, d = self.sizes[i], self.sizes[i - 1]
      x = bufs[f'y{i - 1}']
      if param_grads:
        _linear_wgrad(B, hdim, d, dy, hdim, x.data_ptr(), d, P.g(f'l{i}.w'), P.g(f'l{i}.b'), self)
      dx = gbufs[f'dy{i - 1}'].data_ptr()
      # y_{i-1} is ELU output for i-1 >= 1; for i-1 == 0 it is tanh(LN(.)), handled by the LN backward
      _linear_dgrad(B, hdim, d, dy, hdim, P.p(f'l{i}.w'), dx, d, x.data_ptr() if i - 1 >= 1 else None,
                    ACT_ELU, self)
      dy = dx
    h0 = self.sizes[0]
    _capi.call('b200rl_layernorm_tanh_bwd', B, h0, dy, bufs['y0'].data_ptr(), bufs['xhat'].data_ptr(),
               bufs['rstd'].data_ptr(), P.p('ln.scale'), gbufs['dz0'].data_ptr(),
               P.g('ln.scale') if param_grads else None, P.g('ln.offset') if param_grads else None, st)
    dz0 = gbufs['dz0'].data_ptr()
    if param_grads:
      _linear_wgrad(B, h0, self.in_dim, dz0, h0, x_ptr, self.in_dim, P.g('l0.w'), P.g('l0.b'), self)
    if input_grad:
      _linear_dgrad(B, h0, self.in_dim, dz0, h0, P.p('l0.w'), gbufs['dx'].data_ptr(), self.in_dim, None, ACT_NONE, self)


class D4PGCritic(LayerNormMLP):
  """CriticMultiplexer() -> LayerNormMLP(sizes, activate_final=True) -> DiscreteValuedHead."""

  def __init__(self, obs_dim: int, act_dim: int, sizes=(512, 512, 256), vmin=-150., vmax=150., num_atoms=51,
               device: int = 0, precision: int = _capi.PRECISION_FP32, seed: int = 0):
    super().__init__(obs_dim + act_dim, sizes, num_atoms, 'head', None, device, precision, seed)
    self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
    self.vmin, self.vmax, self.K = float(vmin), float(vmax), int(num_atoms)

  def logits(self, obs, act, bufs):
    """obs f32 [B, obs_dim], act f32 [B, act_dim] -> logits [B, K] (input concat kept in bufs['x'])."""
    _capi.call('b200rl_concat2', bufs['B'], self.obs_dim, self.act_dim, obs.data_ptr(), act.data_ptr(),
               bufs['x'].data_ptr(), _capi.current_stream())
    return self.forward_flat(bufs['x'].data_ptr(), bufs)


class DDPGCritic(LayerNormMLP):
  """CriticMultiplexer(critic_network=LayerNormMLP(sizes + [1])) (`acme/agents/tf/ddpg/agent_test.py:46-47`): the scalar
  critic of DDPG -- the same torso as the D4PG critic with one linear output instead of the 51 atoms."""

  def __init__(self, obs_dim: int, act_dim: int, sizes=(512, 512, 256), device: int = 0,
               precision: int = _capi.PRECISION_FP32, seed: int = 0):
    super().__init__(obs_dim + act_dim, sizes, 1, 'q', None, device, precision, seed)
    self.obs_dim, self.act_dim, self.K = int(obs_dim), int(act_dim), 1

  def logits(self, obs, act, bufs):
    """q(obs, act) [B, 1] (named like D4PGCritic.logits so both critics drive the same learner plumbing)."""
    _capi.call('b200rl_concat2', bufs['B'], self.obs_dim, self.act_dim, obs.data_ptr(), act.data_ptr(),
               bufs['x'].data_ptr(), _capi.current_stream())
    return self.forward_flat(bufs['x'].data_ptr(), bufs)


class D4PGPolicy(LayerNormMLP):
  """LayerNormMLP(sizes, activate_final=True) -> NearZeroInitializedLinear(A) -> TanhToSpec."""

  def __init__(self, obs_dim: int, act_dim: int, sizes=(256, 256, 256), act_min=-1., act_max=1.,
               device: int = 0, precision: int = _capi.PRECISION_FP32, seed: int = 0):
    import torch
    super().__init__(obs_dim, sizes, act_dim, 'out', 1e-4, device, precision, seed)
    self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
    dev = torch.device('cuda', device)
    lo = np.broadcast_to(np.asarray(act_min, np.float32), (act_dim,))
    hi = np.broadcast_to(np.asarray(act_max, np.float32), (act_dim,))
    self.scale = torch.as_tensor(np.ascontiguousarray(hi - lo)).to(dev)
    self.offset = torch.as_tensor(np.ascontiguousarray(lo)).to(dev)

  def make_buffers(self, B: int):
    import torch
    bufs = super().make_buffers(B)
    bufs['a'] = torch.empty((B, self.act_dim), dtype=torch.float32, device=bufs['out'].device)
    bufs['dpre'] = torch.empty_like(bufs['a'])
    return bufs

  def action(self, obs, bufs):
    pre = self.forward_flat(obs.data_ptr(), bufs)
    _capi.call('b200rl_tanh_to_spec_fwd', bufs['B'], self.act_dim, pre.data_ptr(), self.scale.data_ptr(),
               self.offset.data_ptr(), bufs['a'].data_ptr(), _capi.current_stream())
    return bufs['a']

  def backward_action(self, obs, bufs, gbufs, da):
    _capi.call('b200rl_tanh_to_spec_bwd', bufs['B'], self.act_dim, da.data_ptr(), bufs['out'].data_ptr(),
               self.scale.data_ptr(), bufs['dpre'].data_ptr(), _capi.current_stream())
    self.backward_flat(obs.data_ptr(), bufs, gbufs, bufs['dpre'].data_ptr(), param_grads=True, input_grad=False)
