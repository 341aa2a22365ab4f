"""K3 parity (GPU): pooled lookup vs the oracle (bit-exact expected; asserted at
the north-star tolerance 1e-5 relative) for C1-shaped inputs, edge cases, dims."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # BASELINE.json north_star: "within 1e-5 relative on fp32 pooled embeddings"


def _bags(rng, nb, rows, mean_len=3, allow_empty=False):
  lens = rng.poisson(mean_len, nb) + (0 if allow_empty else 1)
  offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
  ids = rng.randint(0, rows, int(offsets[-1])).astype(np.int64)
  return ids, offsets


@pytest.mark.parametrize('combiner', ['mean', 'sum', 'sqrtn'])
def test_c1_config(hb, oracle, combiner):
  """BASELINE configs[0]: 4 tables x 100k rows x dim 16, batch 4096, fp32."""
  rng = np.random.RandomState(0)
  B, D, rows = 4096, 16, 100000
  tables = [rng.uniform(-1e-3, 1e-3, (rows, D)).astype(np.float32) for _ in range(4)]
  feats = [_bags(rng, B, rows) for _ in range(4)]
  gl = hb.embedding.GroupLookup([torch.from_numpy(t).cuda() for t in tables], [combiner] * 4)
  out = gl.forward([torch.from_numpy(f[0]).cuda() for f in feats],
                   [torch.from_numpy(f[1]).cuda() for f in feats], check=True).cpu().numpy()
  assert out.shape == (B, 4 * D)
  for k in range(4):
    exp = oracle.embedding_lookup_sparse(tables[k], feats[k][0], feats[k][1], combiner)
    got = out[:, k * D:(k + 1) * D]
    np.testing.assert_allclose(got, exp, rtol=RTOL, atol=0)
    assert np.array_equal(got, exp), 'expected bit-exact fp32 (same summation order)'


@pytest.mark.parametrize('dim', [4, 8, 16, 32, 48, 64, 128, 256, 516, 1024])
def test_dims(hb, oracle, dim):
  rng = np.random.RandomState(dim)
  rows, B = 5000, 777
  table = rng.randn(rows, dim).astype(np.float32)
  ids, off = _bags(rng, B, rows, 2, allow_empty=True)
  out = hb.embedding.embedding_lookup_sparse(torch.from_numpy(table).cuda(), torch.from_numpy(ids).cuda(),
                                             torch.from_numpy(off).cuda(), combiner='mean', check=True)
  exp = oracle.embedding_lookup_sparse(table, ids, off, 'mean')
  np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=RTOL, atol=0)


def test_one_id_per_bag_and_plain_lookup(hb, oracle):
  rng = np.random.RandomState(1)
  rows, D, B = 100000, 32, 65536
  table = rng.randn(rows, D).astype(np.float32)
  ids = rng.randint(0, rows, B).astype(np.int64)
  t = torch.from_numpy(table).cuda()
  out = hb.embedding.embedding_lookup_sparse(t, torch.from_numpy(ids).cuda(), None, combiner='mean', check=True)
  np.testing.assert_array_equal(out.cpu().numpy(), table[ids])
  out = hb.embedding.embedding_lookup(t, torch.from_numpy(ids).cuda(), check=True)
  np.testing.assert_array_equal(out.cpu().numpy(), table[ids])


def test_empty_bags_and_empty_feature(hb, oracle):
  table = torch.randn(10, 8, device='cuda')
  off = torch.tensor([0, 0, 2, 2, 3], device='cuda')
  ids = torch.tensor([1, 2, 3], device='cuda')
  out = hb.embedding.embedding_lookup_sparse(table, ids, off, combiner='sum', check=True)
  t = table.cpu()
  assert torch.equal(out[0].cpu(), torch.zeros(8)) and torch.equal(out[2].cpu(), torch.zeros(8))
  assert torch.allclose(out[1].cpu(), t[1] + t[2]) and torch.equal(out[3].cpu(), t[3])
  out = hb.embedding.embedding_lookup_sparse(table, torch.empty(0, dtype=torch.int64, device='cuda'),
                                             torch.zeros(1, dtype=torch.int64, device='cuda'))
  assert out.shape == (0, 8)


def test_long_bags_hot_keys(hb, oracle):
  rng = np.random.RandomState(2)
  rows, D, B = 1000, 16, 300
  table = rng.randn(rows, D).astype(np.float32)
  lens = rng.randint(0, 200, B)
  off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
  ids = (rng.zipf(1.2, int(off[-1])) % rows).astype(np.int64)
  for comb in ('sum', 'mean', 'sqrtn'):
    out = hb.embedding.embedding_lookup_sparse(torch.from_numpy(table).cuda(), torch.from_numpy(ids).cuda(),
                                               torch.from_numpy(off).cuda(), combiner=comb, check=True)
    exp = oracle.embedding_lookup_sparse(table, ids, off, comb)
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=RTOL, atol=1e-7)


def test_out_of_range_id_raises(hb):
  table = torch.randn(10, 8, device='cuda')
  with pytest.raises(IndexError):
    hb.embedding.embedding_lookup_sparse(table, torch.tensor([1, 10], device='cuda'), None, check=True)
  with pytest.raises(IndexError):
    hb.embedding.embedding_lookup_sparse(table, torch.tensor([-1], device='cuda'), None, check=True)


def test_criteo_shape_full_size_properties(hb):
  """BASELINE configs[1] shape (26 feats, B=65536, D=32; vocab capped for the
  test's memory): pure-gather property out[b] == table[id[b]] checked on device,
  and linearity of sum pooling."""
  from conftest import criteo_table_sizes
  g = torch.Generator(device='cuda').manual_seed(0)
  sizes = [min(n, 2_000_000) for n in criteo_table_sizes()]
  B, D = 65536, 32
  tables = [torch.randn(n, D, device='cuda', generator=g) for n in sizes]
  ids = [torch.randint(0, n, (B,), device='cuda', generator=g) for n in sizes]
  gl = hb.embedding.GroupLookup(tables, ['mean'] * 26)
  out = gl.forward(ids, check=True)
  for k in range(26):
    assert torch.equal(out[:, k * D:(k + 1) * D], tables[k][ids[k]])
  # linearity: lookup(2*T) == 2*lookup(T) exactly in fp32
  out2 = hb.embedding.embedding_lookup_sparse(tables[0] * 2, ids[0])
  assert torch.equal(out2, 2 * out[:, :D])


def test_offsets_beyond_the_id_buffer_are_flagged(hb):
  """A bag reaching past nnz raises BAD_OFFSETS instead of reading past the ids
  (ADVICE round 1)."""
  t = torch.randn(50, 8, device='cuda')
  ids = torch.arange(10, device='cuda')
  off = torch.tensor([0, 4, 25], device='cuda')  # second bag claims ids[4:25], nnz is 10
  hb.embedding.embedding_lookup_sparse(t, ids, off, combiner='sum')
  with pytest.raises(ValueError):
    hb._util.check_status(t.device)
