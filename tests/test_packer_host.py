"""Host-side packing of steps for the C layer (`acme_b200/replay.py::_Packer`, `Writer.append_step`): CPU only."""
import numpy as np
import pytest

from acme_b200 import replay, specs


def _reference_row(packer, value_nest):
  from acme_b200 import tree
  row = np.zeros(max(packer.nbytes, 1), np.uint8)
  for v, (shape, dt, off, n) in zip(tree.flatten(value_nest), packer.leaves):
    row[off:off + n] = np.frombuffer(np.asarray(v, dtype=dt).tobytes(), np.uint8)
  return row


@pytest.mark.parametrize('spec,value', [
    (specs.Array((84, 84, 4), np.uint8), np.random.default_rng(0).integers(0, 256, (84, 84, 4), dtype=np.uint8)),
    (specs.Array((67,), np.float32), np.random.default_rng(1).standard_normal(67).astype(np.float32)),
    (specs.DiscreteArray(18), np.int32(7)),
    (specs.DiscreteArray(18), 7),
    (specs.Array((), np.float32), 0.25),
    (specs.Array((3, 5), np.float32), np.asfortranarray(np.arange(15, dtype=np.float32).reshape(3, 5))),   # not C-contiguous
    (specs.Array((3, 5), np.float32), np.arange(15, dtype=np.float64).reshape(3, 5)),                       # converts dtype
])
def test_single_leaf_pack_matches_byte_copy(spec, value):
  p = replay._Packer(spec)
  assert p.single
  got = p.pack(value)
  assert got.dtype == np.uint8 and got.flags.c_contiguous
  np.testing.assert_array_equal(got, _reference_row(p, value))


def test_nested_pack_and_shape_errors():
  nest = {'a': specs.Array((2,), np.float32), 'b': specs.Array((3,), np.uint8)}
  p = replay._Packer(nest)
  assert not p.single
  v = {'a': np.array([1., 2.], np.float32), 'b': np.array([1, 2, 3], np.uint8)}
  np.testing.assert_array_equal(p.pack(v), _reference_row(p, v))
  with pytest.raises(ValueError):
    replay._Packer(specs.Array((2, 2), np.float32)).pack(np.zeros((2, 3), np.float32))


def test_writer_sends_the_first_observation_once(monkeypatch):
  """The ring stores each observation once: only an episode's first append carries `observation`."""
  calls = []

  def fake_call(name, *args):
    calls.append((name, args))
    return 0
  monkeypatch.setattr(replay._capi, 'call', fake_call)

  class FakeTable:
    name, handle, _handle, has_extras = 'priority_table', 1, 1, False
    obs_packer = replay._Packer(specs.Array((4,), np.uint8))
    act_packer = replay._Packer(specs.DiscreteArray(3))

  class FakeServer:
    tables = {'priority_table': FakeTable()}

  class FakeClient:
    server = FakeServer()
  w = replay.Writer(FakeClient(), 1)
  o = [np.full(4, i, np.uint8) for i in range(4)]
  for i in range(3):
    w.append_step(o[i], np.int32(i), 0.5, 1.0, o[i + 1])
  appends = [a for n, a in calls if n == 'b200rl_writer_append']
  assert len(appends) == 3
  assert appends[0][2] is not None and appends[1][2] is None and appends[2][2] is None   # obs pointer
  assert all(a[6] is not None for a in appends)                                          # next_obs pointer


@pytest.mark.parametrize('nest', [
    {'a': specs.Array((3,), np.float32), 'b': specs.Array((1,), np.uint8)},          # 13 payload bytes
    (specs.Array((), np.float64), specs.Array((), np.int32)),                        # 12 payload bytes
    (specs.Array((5,), np.uint8), specs.Array((2,), np.float64), specs.Array((3,), np.int16)),
])
def test_mixed_dtype_rows_stay_aligned_in_a_batch(nest):
  """the ROW size must be a multiple of the largest leaf alignment, or `rows[:, off:off+n].view(dtype)`
  fails (and rows b > 0 are misaligned on the device too).  Pack B rows, unpack them as a batch."""
  import torch
  from acme_b200 import tree
  p = replay._Packer(nest)
  assert p.nbytes % max(min(dt.itemsize, 16) for _, dt, _, _ in p.leaves) == 0
  rng = np.random.default_rng(5)
  B = 4
  values = []
  for _ in range(B):
    leaves = [(rng.standard_normal(shape) * 50).astype(dt) for shape, dt, _, _ in p.leaves]
    values.append(tree.unflatten_as(nest, leaves))
  rows = torch.as_tensor(np.stack([p.pack(v) for v in values]))
  assert rows.shape == (B, p.nbytes)
  out = tree.flatten(p.unpack_batch(rows))
  for j, (shape, dt, _, _) in enumerate(p.leaves):
    want = np.stack([np.asarray(tree.flatten(v)[j], dt) for v in values])
    np.testing.assert_array_equal(out[j].numpy(), want)
