"""K1 parity (GPU, through the C-ABI): bit-exact vs the oracle and vs the
reference-generated fixtures; reference-style property tests; edge cases."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

_TORCH = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}


def _to_dev(x):
  if x.dtype in _TORCH:
    return torch.from_numpy(x).cuda()
  # uint32/uint64: move the bits, then reinterpret
  signed = x.view(np.int32 if x.dtype == np.uint32 else np.int64)
  t = torch.from_numpy(signed.copy()).cuda()
  return t.view(torch.uint32 if x.dtype == np.uint32 else torch.uint64)


def _to_host(t, like):
  if like.dtype in _TORCH:
    return t.cpu().numpy()
  s = t.view(torch.int32 if like.dtype == np.uint32 else torch.int64).cpu().numpy()
  return s.view(like.dtype)


def test_partition_fixture_bit_exact(hb, golden_partition):
  g = golden_partition
  names = sorted({k.split('/')[1] for k in g if k.startswith('mod/')})
  for name in names:
    x, p = g[f'mod/{name}/x'], int(g[f'mod/{name}/p'])
    y, s, i = hb.distribute.partition_by_modulo(_to_dev(x), p)
    np.testing.assert_array_equal(_to_host(y, x), g[f'mod/{name}/y'], err_msg=name)
    np.testing.assert_array_equal(s.cpu().numpy(), g[f'mod/{name}/sizes'], err_msg=name)
    np.testing.assert_array_equal(i.cpu().numpy(), g[f'mod/{name}/idx'], err_msg=name)


def test_dual_partition_fixture_bit_exact(hb, golden_partition):
  g = golden_partition
  names = sorted({k.split('/')[1] for k in g if k.startswith('dual/')})
  for name in names:
    for st in (1, 2):
      key = f'dual/{name}/s{st}'
      p, m, stage = [int(v) for v in g[f'{key}/pm']]
      x = g[f'{key}/x']
      fn = (hb.distribute.partition_by_dual_modulo_stage_one if stage == 1
            else hb.distribute.partition_by_dual_modulo_stage_two)
      y, s, i = fn(_to_dev(x), p, m)
      np.testing.assert_array_equal(_to_host(y, x), g[f'{key}/y'])
      np.testing.assert_array_equal(s.cpu().numpy(), g[f'{key}/sizes'])
      np.testing.assert_array_equal(i.cpu().numpy(), g[f'{key}/idx'])


@pytest.mark.parametrize('dt', [np.int32, np.int64, np.uint32, np.uint64])
@pytest.mark.parametrize('p', [1, 2, 3, 5, 8, 64, 100])
def test_partition_vs_oracle(hb, oracle, dt, p):
  rng = np.random.RandomState(p)
  for n in (1, 31, 2048, 2049, 50000):
    if np.issubdtype(dt, np.signedinteger):
      x = rng.randint(-10**9, 10**9, size=n).astype(dt)
    else:
      x = (rng.randint(0, 2**31, size=n).astype(np.int64) * 3 + 1).astype(dt)
    y, s, i = hb.distribute.partition_by_modulo(_to_dev(x), p)
    ey, es, ei = oracle.partition_by_modulo(x, p)
    np.testing.assert_array_equal(_to_host(y, x), ey)
    np.testing.assert_array_equal(s.cpu().numpy(), es)
    np.testing.assert_array_equal(i.cpu().numpy(), ei)


def test_unfused_reference_style(hb):
  # partition_test.py:40-65
  np.random.seed(0)
  x = np.random.randint(low=-1000000000, high=1000000000, size=10000, dtype=np.int32)
  y, ysizes, idx = hb.distribute.partition_by_modulo(torch.from_numpy(x).cuda(), 5)
  y, ysizes, idx = y.cpu().numpy(), ysizes.cpu().numpy(), idx.cpu().numpy()
  assert len(y) == len(idx) and len(ysizes) == 5
  np.testing.assert_array_equal(x, np.take(y, idx))


def test_empty_input(hb):
  # partition_test.py:67-81
  y, ysizes, idx = hb.distribute.partition_by_modulo(torch.empty(0, dtype=torch.int64, device='cuda'), 7)
  assert y.numel() == 0 and idx.numel() == 0 and ysizes.tolist() == [0] * 7


def test_fused_n(hb, oracle):
  # partition_test.py:83-114: 10 columns x 100000 int64, P=3, one packed call
  np.random.seed(0)
  xs = [np.random.randint(low=-1000000000, high=1000000000, size=100000, dtype=np.int64)
        for _ in range(10)]
  ys, ss, iis = hb.distribute.partition_by_modulo([torch.from_numpy(x).cuda() for x in xs], 3)
  for c in range(10):
    ey, es, ei = oracle.partition_by_modulo(xs[c], 3)
    np.testing.assert_array_equal(ys[c].cpu().numpy(), ey)
    np.testing.assert_array_equal(ss[c].cpu().numpy(), es)
    np.testing.assert_array_equal(iis[c].cpu().numpy(), ei)
    np.testing.assert_array_equal(xs[c], np.take(ys[c].cpu().numpy(), iis[c].cpu().numpy()))


def test_empty_input_fused_and_ragged(hb, oracle):
  # partition_test.py:116-140 + ragged lengths incl. > 128 inputs (chunked launches)
  xs = [torch.empty(0, dtype=torch.int64, device='cuda') for _ in range(3)]
  ys, ss, iis = hb.distribute.partition_by_modulo(xs, 7)
  for c in range(3):
    assert ys[c].numel() == 0 and ss[c].tolist() == [0] * 7
  rng = np.random.RandomState(3)
  lens = [0, 1, 5000, 0, 2048, 4097] + [rng.randint(0, 3000) for _ in range(200)]
  xs = [rng.randint(0, 10**6, size=n).astype(np.int64) for n in lens]
  ys, ss, iis = hb.distribute.partition_by_modulo([torch.from_numpy(x).cuda() for x in xs], 8)
  for c, x in enumerate(xs):
    ey, es, ei = oracle.partition_by_modulo(x, 8)
    np.testing.assert_array_equal(ys[c].cpu().numpy(), ey)
    np.testing.assert_array_equal(ss[c].cpu().numpy(), es)
    np.testing.assert_array_equal(iis[c].cpu().numpy(), ei)


def test_full_size_properties(hb):
  """BASELINE config sizes (26 x 65536 ids, P=8): size-independent properties."""
  rng = np.random.RandomState(0)
  xs = [torch.from_numpy(rng.randint(0, 4 * 10**7, 65536).astype(np.int64)).cuda() for _ in range(26)]
  ys, ss, iis = hb.distribute.partition_by_modulo(xs, 8)
  for x, y, s, i in zip(xs, ys, ss, iis):
    assert int(s.sum()) == x.numel()
    assert torch.equal(y[i.long()], x)                      # inverse permutation
    shard = (y % 8).cpu()
    assert bool((shard[1:] >= shard[:-1]).all())            # grouped by shard
    starts = torch.cumsum(s, 0) - s
    for b in range(8):                                      # stable inside a bucket
      pos = torch.nonzero((x % 8) == b).flatten()
      seg = i[pos].long()
      assert bool((seg[1:] > seg[:-1]).all()) if seg.numel() > 1 else True
      assert seg.numel() == 0 or int(seg[0]) == int(starts[b])


def test_errors(hb):
  with pytest.raises(ValueError, match='1D'):
    hb.distribute.partition_by_modulo(torch.zeros(2, 2, dtype=torch.int64, device='cuda'), 2)
  with pytest.raises(ValueError, match='num_partitions'):
    hb.distribute.partition_by_modulo(torch.zeros(2, dtype=torch.int64, device='cuda'), 0)
  with pytest.raises(TypeError):
    hb.distribute.partition_by_modulo(torch.zeros(2, device='cuda'), 2)
