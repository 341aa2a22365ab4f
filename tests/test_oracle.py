"""The CPU oracle pinned against the reference's golden vectors / own functors.
(-m "not gpu": runs in the survey container and on the GPU box alike.)"""
import numpy as np
import pytest
import torch


def _cases(g, prefix):
  names = sorted({k.split('/')[1] for k in g if k.startswith(prefix + '/')})
  return names


def test_partition_matches_reference_fixture(oracle, golden_partition):
  g = golden_partition
  for name in _cases(g, 'mod'):
    x, p = g[f'mod/{name}/x'], int(g[f'mod/{name}/p'])
    y, s, i = oracle.partition_by_modulo(x, p)
    np.testing.assert_array_equal(y, g[f'mod/{name}/y'], err_msg=name)
    np.testing.assert_array_equal(s, g[f'mod/{name}/sizes'], err_msg=name)
    np.testing.assert_array_equal(i, g[f'mod/{name}/idx'], err_msg=name)


def test_dual_partition_matches_reference_fixture(oracle, golden_partition):
  g = golden_partition
  for name in _cases(g, 'dual'):
    for st in (1, 2):
      key = f'dual/{name}/s{st}'
      p, m, stage = [int(v) for v in g[f'{key}/pm']]
      y, s, i = oracle.partition_by_dual_modulo(g[f'{key}/x'], p, m, stage)
      np.testing.assert_array_equal(y, g[f'{key}/y'])
      np.testing.assert_array_equal(s, g[f'{key}/sizes'])
      np.testing.assert_array_equal(i, g[f'{key}/idx'])


def test_partition_matches_compiled_reference(oracle):
  """Bit-for-bit against the reference's own functor (oracle/_ref), when built."""
  if oracle.ref() is None:
    pytest.skip('oracle/_ref not built (no /root/reference here)')
  rng = np.random.RandomState(1)
  for dt in (np.int32, np.int64, np.uint32, np.uint64):
    lo, hi = (-10**9, 10**9) if np.issubdtype(dt, np.signedinteger) else (0, 2**31)
    x = rng.randint(lo, hi, size=20000).astype(dt)
    for p in (1, 2, 5, 8, 13):
      for a, b in zip(oracle.partition_by_modulo(x, p), oracle.ref_partition_by_modulo(x, p)):
        np.testing.assert_array_equal(a, b)
      for st in (1, 2):
        for a, b in zip(oracle.partition_by_dual_modulo(x, p, 3, st),
                        oracle.ref_partition_by_dual_modulo(x, p, 3, st)):
          np.testing.assert_array_equal(a, b)


def test_partition_properties_reference_style(oracle):
  # partition_test.py:40-65 / :67-81
  np.random.seed(0)
  x = np.random.randint(low=-1000000000, high=1000000000, size=10000, dtype=np.int32)
  y, sizes, idx = oracle.partition_by_modulo(x, 5)
  assert len(y) == len(idx) and len(sizes) == 5
  np.testing.assert_array_equal(x, np.take(y, idx))
  assert sizes.sum() == len(x)
  y, sizes, idx = oracle.partition_by_modulo(np.array([], np.int64), 7)
  assert len(y) == 0 and len(idx) == 0 and (sizes == 0).all() and len(sizes) == 7


def test_murmur3_matches_reference_fixture(oracle, golden_partition):
  for k, h in zip(golden_partition['murmur/keys'], golden_partition['murmur/hash']):
    assert oracle.murmur3_hash32(int(k)) == int(h)


def test_alltoallv_golden(oracle, golden_alltoall):
  g = golden_alltoall['alltoallv']  # alltoall_test.py:219-226
  outs, osz = oracle.alltoallv([np.array(v, np.int64) for v in g['ids']], g['sizes'])
  for r in range(2):
    np.testing.assert_array_equal(outs[r], g['out_ids'][r])
    np.testing.assert_array_equal(osz[r], g['out_sizes'][r])


def test_alltoallv_n_golden(oracle, golden_alltoall):
  g = golden_alltoall['alltoallv_n']  # alltoall_test.py:254-269
  n_in = [[np.array(t['ids'], np.float32) for t in g['inputs'][str(r)]] for r in range(2)]
  n_sz = [[t['sizes'] for t in g['inputs'][str(r)]] for r in range(2)]
  res = oracle.alltoallv_n(n_in, n_sz)
  for r in range(2):
    for k in range(2):
      np.testing.assert_allclose(res[r][k][0], g['outputs'][str(r)][k]['ids'], rtol=1e-6)
      np.testing.assert_array_equal(res[r][k][1], g['outputs'][str(r)][k]['sizes'])


def test_alltoallv_grad_golden(oracle, golden_alltoall):
  # alltoall_test.py:228-243: loss=mean(outputs); the gradient is the reverse
  # alltoallv (with the exchanged sizes) of g/len(outputs) -> g/(sum recv) per elem
  g = golden_alltoall['alltoallv_grad']
  sizes = g['sizes']
  vals = [np.full(sum(s), 1.0, np.float32) for s in sizes]
  outs, osz = oracle.alltoallv(vals, sizes)
  up = [np.full(o.shape, g['g'] / o.shape[0], np.float32) for o in outs]
  back, _ = oracle.alltoallv(up, osz)
  g0 = g['g'] / (sizes[0][0] + sizes[1][0])
  g1 = g['g'] / (sizes[0][1] + sizes[1][1])
  np.testing.assert_allclose(back[0], sizes[0][0] * [g0] + sizes[0][1] * [g1], rtol=1e-6)
  np.testing.assert_allclose(back[1], sizes[1][0] * [g0] + sizes[1][1] * [g1], rtol=1e-6)


def test_alltoallv_common_shape(oracle):
  rng = np.random.RandomState(0)
  W, D = 3, 4
  sizes = rng.randint(0, 5, size=(W, W))
  ins = [rng.randn(int(sizes[r].sum()), D).astype(np.float32) for r in range(W)]
  outs, osz = oracle.alltoallv(ins, sizes, (D,))
  for r in range(W):
    exp = []
    for q in range(W):
      o = int(sizes[q][:r].sum())
      exp.append(ins[q][o:o + sizes[q][r]])
    np.testing.assert_array_equal(outs[r], np.concatenate(exp))
    np.testing.assert_array_equal(osz[r], sizes[:, r])


def test_shard_rule(oracle):
  # embedding/variables.py:95-124
  for n in (1, 7, 8, 9, 1000, 39884406):
    for w in (1, 2, 8):
      rows = [oracle.shard_rows(n, w, s) for s in range(w)]
      assert sum(rows) == n
      assert rows == [len(range(s, n, w)) for s in range(w)]
      offs = [oracle.shard_offset(n, w, s) for s in range(w)]
      assert offs == [sum(rows[:s]) for s in range(w)]
  assert oracle.is_small_table(8, 8) and not oracle.is_small_table(9, 8)
  assert oracle.is_small_table(65536, 8, 65536) and not oracle.is_small_table(65537, 8, 65536)


def test_unique_first_occurrence(oracle):
  u, inv = oracle.unique(np.array([5, 3, 5, 9, 3, 3, 7], np.int64))
  np.testing.assert_array_equal(u, [5, 3, 9, 7])
  np.testing.assert_array_equal(inv, [0, 1, 0, 2, 1, 1, 3])


def _bags(rng, nb, rows, mean_len=3):
  lens = rng.poisson(mean_len, nb) + 1
  offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
  ids = rng.randint(0, rows, int(offsets[-1])).astype(np.int64)
  return ids, offsets


@pytest.mark.parametrize('combiner', ['sum', 'mean'])
def test_lookup_sparse_vs_torch_embedding_bag(oracle, combiner):
  """Cross-check of the PARITY-UNPINNED part against an independent implementation."""
  rng = np.random.RandomState(0)
  rows, dim, nb = 1000, 16, 512
  table = rng.uniform(-1e-3, 1e-3, (rows, dim)).astype(np.float32)
  ids, offsets = _bags(rng, nb, rows)
  out = oracle.embedding_lookup_sparse(table, ids, offsets, combiner)
  ref = torch.nn.functional.embedding_bag(
      torch.from_numpy(ids), torch.from_numpy(table), torch.from_numpy(offsets[:-1]),
      mode=combiner).numpy()
  np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-9)
  out2 = oracle.embedding_lookup_sparse(table, ids, offsets, combiner, dedup=False)
  np.testing.assert_array_equal(out, out2)


def test_lookup_sparse_sqrtn_and_empty_bags(oracle):
  rng = np.random.RandomState(1)
  table = rng.randn(50, 8).astype(np.float32)
  offsets = np.array([0, 2, 2, 5, 5], np.int64)
  ids = np.array([1, 2, 3, 3, 4], np.int64)
  out = oracle.embedding_lookup_sparse(table, ids, offsets, 'sqrtn')
  np.testing.assert_allclose(out[0], (table[1] + table[2]) / np.sqrt(np.float32(2)), rtol=1e-6)
  assert (out[1] == 0).all() and (out[3] == 0).all()
  np.testing.assert_allclose(out[2], (table[3] + table[3] + table[4]) / np.sqrt(np.float32(3)), rtol=1e-6)


def test_sharded_lookup_equals_unsharded(oracle):
  """Invariant sharded(W) == unsharded (embedding/sharding.py:171-203 recipe)."""
  rng = np.random.RandomState(2)
  N, D = 1003, 8
  table = rng.randn(N, D).astype(np.float32)
  for W in (1, 2, 4, 8):
    shards = oracle.shard_table(table, W)
    ids = [rng.randint(0, N, rng.randint(0, 200)).astype(np.int64) for _ in range(W)]
    outs = oracle.sharded_embedding_lookup(shards, N, ids)
    for r in range(W):
      np.testing.assert_array_equal(outs[r], table[ids[r]])


def test_adagrad_closed_form(oracle):
  rng = np.random.RandomState(3)
  rows, dim = 20, 4
  w = rng.randn(rows, dim).astype(np.float32)
  acc = np.full((rows, dim), 0.1, np.float32)
  idx = np.array([3, 5, 3, 7], np.int64)
  g = rng.randn(4, dim).astype(np.float32)
  w0, a0 = w.copy(), acc.copy()
  oracle.sparse_apply_adagrad(w, acc, idx, g, 0.01)
  gs = {3: g[0] + g[2], 5: g[1], 7: g[3]}
  for r in range(rows):
    if r in gs:
      a = a0[r] + gs[r] * gs[r]
      np.testing.assert_allclose(acc[r], a, rtol=1e-6)
      np.testing.assert_allclose(w[r], w0[r] - 0.01 * gs[r] / np.sqrt(a), rtol=1e-6)
    else:
      np.testing.assert_array_equal(w[r], w0[r])


def test_adagrad_vs_torch_single_step(oracle):
  rng = np.random.RandomState(4)
  rows, dim = 30, 8
  w = rng.randn(rows, dim).astype(np.float32)
  acc = np.full((rows, dim), 0.1, np.float32)
  idx = rng.randint(0, rows, 40).astype(np.int64)
  g = rng.randn(40, dim).astype(np.float32)
  p = torch.nn.Parameter(torch.from_numpy(w.copy()))
  opt = torch.optim.Adagrad([p], lr=0.01, initial_accumulator_value=0.1, eps=0.0)
  dense = torch.zeros(rows, dim).index_add_(0, torch.from_numpy(idx), torch.from_numpy(g))
  p.grad = dense
  opt.step()
  oracle.sparse_apply_adagrad(w, acc, idx, g, 0.01)
  touched = np.unique(idx)
  np.testing.assert_allclose(w[touched], p.detach().numpy()[touched], rtol=2e-5, atol=1e-7)


def test_lazy_adam_closed_form(oracle):
  rng = np.random.RandomState(5)
  rows, dim = 10, 4
  w = rng.randn(rows, dim).astype(np.float32)
  m = np.zeros((rows, dim), np.float32)
  v = np.zeros((rows, dim), np.float32)
  idx = np.array([1, 1, 4], np.int64)
  g = rng.randn(3, dim).astype(np.float32)
  w0 = w.copy()
  oracle.sparse_apply_lazy_adam(w, m, v, idx, g, 0.001, 0.9, 0.999, 1e-8, step=1)
  gs = {1: g[0] + g[1], 4: g[2]}
  lr_t = 0.001 * np.sqrt(1 - 0.999) / (1 - 0.9)
  for r, gg in gs.items():
    mm = 0.1 * gg
    vv = 0.001 * gg * gg
    np.testing.assert_allclose(m[r], mm, rtol=1e-5)
    np.testing.assert_allclose(w[r], w0[r] - lr_t * mm / (np.sqrt(vv) + 1e-8), rtol=1e-4)
  assert (m[0] == 0).all() and (w[0] == w0[0]).all()


def test_row_grads(oracle):
  grad = np.arange(12, dtype=np.float32).reshape(3, 4)
  offsets = np.array([0, 2, 2, 5], np.int64)
  rg = oracle.lookup_row_grads(grad, offsets, 'mean')
  np.testing.assert_allclose(rg[0], grad[0] / 2)
  np.testing.assert_allclose(rg[4], grad[2] / 3)
  rg = oracle.lookup_row_grads(grad, offsets, 'sum')
  np.testing.assert_array_equal(rg[1], grad[0])


def test_cache_lookup_oracle(oracle):
  slabs = 7
  cache = np.full(slabs * 32, np.iinfo(np.int64).min, np.int64)
  present = np.arange(100, 160, dtype=np.int64)
  for k in present:  # insert with the reference's probing rule
    s = oracle.murmur3_hash32(int(k)) % slabs
    while True:
      row = cache[s * 32:(s + 1) * 32]
      free = np.where(row == np.iinfo(np.int64).min)[0]
      if len(free):
        row[free[0]] = k
        break
      s = (s + 1) % slabs
  keys = np.array([100, 5, 159, 777, 130], np.int64)
  hi, hc, mi, mk = oracle.cache_lookup(cache, keys)
  np.testing.assert_array_equal(hi, [0, 2, 4])
  np.testing.assert_array_equal(cache[hc], [100, 159, 130])
  np.testing.assert_array_equal(mi, [1, 3])
  np.testing.assert_array_equal(mk, [5, 777])


def test_mt_step_equals_single_thread(oracle):
  """oracle/hb_oracle_mt.c (pthread driver of bench.py's CPU legs) == the plain oracle."""
  rng = np.random.RandomState(0)
  F, B, D = 5, 3000, 8
  rows = [50000, 3, 700, 2000000, 64]
  T = [rng.randn(r, D).astype(np.float32) for r in rows]
  A = [np.full((r, D), 0.1, np.float32) for r in rows]
  T2 = [t.copy() for t in T]
  A2 = [a.copy() for a in A]
  ids = [(rng.zipf(1.2, B) % r).astype(np.int64) for r in rows]
  grad = rng.randn(B, F * D).astype(np.float32)
  out = np.empty((B, F * D), np.float32)
  out2 = np.empty_like(out)
  oracle.mt_step(T, A, ids, grad, out, 0.01, 8, fwd_chunk=777, parts=[4, 1, 2, 8, 1])
  off = np.arange(B + 1, dtype=np.int64)
  for k in range(F):
    oracle.embedding_lookup_sparse(T2[k], ids[k], off, 'mean', out=out2[:, k * D:(k + 1) * D])
  for k in range(F):
    oracle.sparse_apply_adagrad(T2[k], A2[k], ids[k], np.ascontiguousarray(grad[:, k * D:(k + 1) * D]), 0.01)
  assert np.array_equal(out, out2)
  for k in range(F):
    assert np.array_equal(T[k], T2[k]) and np.array_equal(A[k], A2[k])
