"""Property tests (hypothesis) of the CPU oracle: the invariants the reference's own
tests check (partition_test.py:59, :114: `take(y, idx) == x`, sizes) plus stability,
alltoallv inverse, unique/segment identities -- on ragged and degenerate inputs."""
import numpy as np
from hypothesis import given, settings, strategies as st


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(-2**62, 2**62), max_size=300), st.integers(1, 40))
def test_partition_properties(oracle, xs, p):
  x = np.asarray(xs, np.int64)
  y, sizes, idx = oracle.partition_by_modulo(x, p)
  assert len(sizes) == p and sizes.sum() == len(x)
  assert np.array_equal(np.take(y, idx), x) if len(x) else len(y) == 0
  shard = np.mod(y, p)                      # floor modulo == (v % p + p) % p
  assert np.all(np.diff(shard) >= 0)        # grouped by shard, ascending
  assert np.array_equal(np.bincount(shard.astype(np.int64), minlength=p)[:p], sizes)
  for b in range(p):                        # stable inside a bucket
    assert np.all(np.diff(idx[np.mod(x, p) == b]) > 0)


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(0, 2**32 - 1), max_size=200), st.integers(1, 6), st.integers(1, 6),
       st.sampled_from([1, 2]))
def test_dual_partition_properties(oracle, xs, p, m, stage):
  x = np.asarray(xs, np.uint32)
  y, sizes, idx = oracle.partition_by_dual_modulo(x, p, m, stage)
  assert sizes.sum() == len(x)
  if len(x):
    assert np.array_equal(np.take(y, idx), x)
    pre = x.astype(np.int64) % (p * m)
    shard = pre % p if stage == 1 else pre // m
    assert np.array_equal(np.bincount(shard, minlength=p)[:p], sizes)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 5), st.integers(0, 3), st.data())
def test_alltoallv_roundtrip(oracle, w, d, data):
  sizes = np.asarray(data.draw(st.lists(st.lists(st.integers(0, 6), min_size=w, max_size=w),
                                        min_size=w, max_size=w)), np.int32)
  shape = (d,) if d else ()
  rng = np.random.RandomState(0)
  ins = [rng.randint(0, 1000, (int(sizes[r].sum()),) + shape).astype(np.int64) for r in range(w)]
  outs, osz = oracle.alltoallv(ins, sizes, shape)
  for r in range(w):
    assert np.array_equal(osz[r], sizes[:, r])
  back, bsz = oracle.alltoallv(outs, osz, shape)   # the backward exchange restores the layout
  for r in range(w):
    assert np.array_equal(back[r], ins[r]) and np.array_equal(bsz[r], sizes[r])


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(0, 50), max_size=200))
def test_unique_properties(oracle, xs):
  x = np.asarray(xs, np.int64)
  u, inv = oracle.unique(x)
  assert len(set(u.tolist())) == len(u)
  assert np.array_equal(u[inv], x) if len(x) else len(u) == 0
  first = [np.flatnonzero(x == v)[0] for v in u]
  assert first == sorted(first)             # first-occurrence order (tf.unique)


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 64), st.integers(1, 8), st.integers(1, 40), st.sampled_from(['sum', 'mean', 'sqrtn']),
       st.integers(0, 2**31))
def test_lookup_sparse_linear_and_dedup_invariant(oracle, rows, dq, nbags, comb, seed):
  rng = np.random.RandomState(seed)
  dim = 4 * dq
  table = rng.randn(rows, dim).astype(np.float32)
  lens = rng.randint(0, 5, nbags)
  off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
  ids = rng.randint(0, rows, int(off[-1])).astype(np.int64)
  a = oracle.embedding_lookup_sparse(table, ids, off, comb)
  b = oracle.embedding_lookup_sparse(table, ids, off, comb, dedup=False)
  assert np.array_equal(a, b)               # unique->gather->segment == direct bag sum
  assert np.all(a[lens == 0] == 0)          # empty bags are zero rows
  two = oracle.embedding_lookup_sparse(2 * table, ids, off, comb)
  assert np.array_equal(two, 2 * a)         # exact linearity in fp32 (power of two)
