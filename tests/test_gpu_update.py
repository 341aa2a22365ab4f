"""K5 parity (GPU): backward + sparse optimizer vs the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rank_cases  # noqa: E402  pylint: disable=wrong-import-position

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _bags(rng, nb, rows, mean_len=3):
  lens = rng.poisson(mean_len, nb) + 1
  offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
  ids = rng.randint(0, rows, int(offsets[-1])).astype(np.int64)
  return ids, offsets


def _oracle_step(oracle, opt, tables, slots, feats, grad, cols, combiner, step):
  for k, (ids, off) in enumerate(feats):
    D = tables[k].shape[1]
    g = np.ascontiguousarray(grad[:, cols[k]:cols[k] + D])
    if off is None:
      off = np.arange(len(ids) + 1, dtype=np.int64)
    rg = oracle.lookup_row_grads(g, off, combiner)
    if opt == 'adagrad':
      oracle.sparse_apply_adagrad(tables[k], slots[k][0], ids, rg, 0.01)
    elif opt == 'lazy_adam':
      oracle.sparse_apply_lazy_adam(tables[k], slots[k][0], slots[k][1], ids, rg, 0.001, 0.9, 0.999,
                                    1e-8, step)


def _run(hb, oracle, opt, combiner, rows_list, D, B, steps=2, one_hot=False, zipf=False, seed=0,
         rtol=RTOL):
  rng = np.random.RandomState(seed)
  n = len(rows_list)
  tables = [rng.uniform(-1e-1, 1e-1, (r, D)).astype(np.float32) for r in rows_list]
  dev_tables = [torch.from_numpy(t.copy()).cuda() for t in tables]
  if opt == 'adagrad':
    optimizer = hb.training.Adagrad(0.01)
    slots = [[np.full_like(t, 0.1)] for t in tables]
  else:
    optimizer = hb.training.LazyAdam(0.001)
    slots = [[np.zeros_like(t), np.zeros_like(t)] for t in tables]
  gl = hb.embedding.GroupLookup(dev_tables, [combiner] * n)
  cols = [k * D for k in range(n)]
  for step in range(1, steps + 1):
    feats = []
    for r in rows_list:
      if one_hot:
        ids = (rng.zipf(1.3, B) % r if zipf else rng.randint(0, r, B)).astype(np.int64)
        feats.append((ids, None))
      else:
        feats.append(_bags(rng, B, r))
    grad = rng.randn(B, n * D).astype(np.float32)
    gl.forward([torch.from_numpy(f[0]).cuda() for f in feats],
               None if one_hot else [torch.from_numpy(f[1]).cuda() for f in feats])
    gl.backward_update(torch.from_numpy(grad).cuda(), optimizer, check=True)
    _oracle_step(oracle, opt, tables, slots, feats, grad, cols, combiner, step)
  for k in range(n):
    got = dev_tables[k].cpu().numpy()
    np.testing.assert_allclose(got, tables[k], rtol=rtol, atol=1e-6, err_msg=f'table {k}')
    for s_dev, s_ref in zip(gl.slots(k), slots[k]):
      np.testing.assert_allclose(s_dev.cpu().numpy(), s_ref, rtol=rtol, atol=1e-6)
  return dev_tables, tables


@pytest.mark.parametrize('combiner', ['mean', 'sum', 'sqrtn'])
def test_adagrad_c1(hb, oracle, combiner):
  _run(hb, oracle, 'adagrad', combiner, [100000] * 4, 16, 4096)


def test_adagrad_one_hot_no_dups_bit_exact(hb, oracle):
  # rows >> B: almost no duplicates -> identical op sequence -> bit-exact expected
  dev, ref = _run(hb, oracle, 'adagrad', 'mean', [3000000, 500000], 32, 8192, one_hot=True)
  for d, r in zip(dev, ref):
    assert np.array_equal(d.cpu().numpy(), r)


def _run_bounded(hb, oracle, rows_list, D, B, gen, steps=2, lr=0.01, fast_math=False):
  """Adagrad vs the oracle with the PER-ELEMENT summation-order bound of
  rank_cases.adagrad_reference (a formula of each row's run length and gradient
  magnitudes) instead of a blanket tolerance."""
  rng = np.random.RandomState(0)
  n = len(rows_list)
  tables = [rng.uniform(-1e-1, 1e-1, (r, D)).astype(np.float32) for r in rows_list]
  accs = [np.full_like(t, 0.1) for t in tables]
  dev_tables = [torch.from_numpy(t.copy()).cuda() for t in tables]
  gl = hb.embedding.GroupLookup(dev_tables, ['mean'] * n)
  opt = hb.training.Adagrad(lr, fast_math=fast_math)
  tol = [None] * n
  for step in range(steps):
    ids = [gen(rng, B, r) for r in rows_list]
    grad = rng.randn(B, n * D).astype(np.float32)
    gl.forward([torch.from_numpy(i).cuda() for i in ids])
    gl.backward_update(torch.from_numpy(grad).cuda(), opt, check=True)
    for k in range(n):
      g = np.ascontiguousarray(grad[:, k * D:(k + 1) * D])
      tw, ta = rank_cases.adagrad_reference(oracle, tables[k], accs[k], ids[k], g, lr, tol[k])
      if fast_math:  # + 8 ulp of the largest step (sqrt.approx + div.approx) per optimizer step
        tw = tw + 8 * rank_cases.U * lr * np.abs(g).max() * 50 / np.sqrt(0.1)
      tol[k] = (tw, ta)
      got = dev_tables[k].cpu().numpy()
      bad = np.abs(got.astype(np.float64) - tables[k]) > tw
      assert not bad.any(), (f'step {step} table {k} ({rows_list[k]} rows): {int(bad.sum())} elements beyond '
                             f'the bound; max |diff| {np.abs(got - tables[k]).max():.3e}')
      gacc = gl.slots(k)[0].cpu().numpy()
      assert not (np.abs(gacc.astype(np.float64) - accs[k]) > ta).any(), f'step {step} accumulator {k}'
  return dev_tables, tables


def test_adagrad_hot_rows_tiny_tables(hb, oracle):
  # Criteo has tables of 3..155 rows: thousands of duplicates per row -> every row
  # is a multi-piece run of the long kernel.  The oracle adds a row's gradients
  # strictly left to right, the kernels in a fixed piece tree: both are valid fp32
  # sums of the same terms; the bound is a formula of the run length.
  _run_bounded(hb, oracle, [3, 4, 10, 63, 155, 976], 32, 20000, lambda rng, B, r: rng.randint(0, r, B).astype(np.int64))


def test_adagrad_zipf(hb, oracle):
  _run_bounded(hb, oracle, [1543, 39043, 403346], 32, 30000,
               lambda rng, B, r: (rng.zipf(1.3, B) % r).astype(np.int64))


def test_adagrad_run_lengths_around_the_kernel_thresholds(hb, oracle):
  """Runs of exactly 1, 2, 3, 16, 17 (short/long switch), 255, 256, 257 (piece
  boundary), 512, 513 and 5000 entries, interleaved in one feature."""
  lens = [1, 2, 3, 15, 16, 17, 18, 255, 256, 257, 511, 512, 513, 5000, 1, 2]

  def gen(rng, B, r):
    ids = np.concatenate([np.full(n, 7 + 3 * i, np.int64) for i, n in enumerate(lens)])
    rest = rng.randint(1000, r, B - len(ids)).astype(np.int64)
    ids = np.concatenate([ids, rest])
    rng.shuffle(ids)
    return ids
  _run_bounded(hb, oracle, [200000], 32, 12000, gen, steps=1)
  _run_bounded(hb, oracle, [200000], 64, 12000, gen, steps=1)
  _run_bounded(hb, oracle, [200000], 16, 12000, gen, steps=1)
  _run_bounded(hb, oracle, [200000], 256, 12000, gen, steps=1)


def test_adagrad_more_than_eight_tiles_per_feature(hb, oracle):
  """150 000 ids in ONE feature = 19 tiles of the cluster sort: every CTA of the cluster
  owns 2-3 tiles and re-ranks them in the scatter phase (the one-tile-per-CTA fast path
  keeps ranks in registers); 3 digit passes (500 k rows) and 1 pass (300 rows)."""
  _run_bounded(hb, oracle, [500000, 300], 16, 150000,
               lambda rng, B, r: (rng.zipf(1.2, B) % r).astype(np.int64), steps=1)


def test_adagrad_fast_math_within_bound(hb, oracle):
  """HB_OPT_FLAG_FAST_MATH (sqrt.approx / div.approx, the arithmetic class of TF's
  GPU kernels): within 1e-5 relative of the IEEE result, and inside the run-length
  bound plus 8 ulp of the step."""
  dev, ref = _run_bounded(hb, oracle, [100000, 300], 32, 8192,
                          lambda rng, B, r: rng.randint(0, r, B).astype(np.int64), fast_math=True)
  for d, r in zip(dev, ref):
    np.testing.assert_allclose(d.cpu().numpy(), r, rtol=1e-5, atol=1e-6)


def test_c2_config_as_benched(hb, oracle):
  """BASELINE configs[1] exactly as bench.py runs it: the 26 Criteo-Terabyte table
  sizes (capped at 2M rows so the oracle's host tables stay small), Zipf(1.05) ids
  from bench.gen_ids_numpy, batch 65 536, dim 32, mean combiner, two steps of
  forward + backward + Adagrad(0.01), against the oracle."""
  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  import bench
  sizes = [min(n, 2000000) for n in bench.CRITEO_SIZES]
  B, D, lr = 65536, 32, 0.01
  rng = np.random.RandomState(1234)
  tables = [rng.uniform(-1e-3, 1e-3, (r, D)).astype(np.float32) for r in sizes]
  accs = [np.full_like(t, 0.1) for t in tables]
  dev_tables = [torch.from_numpy(t.copy()).cuda() for t in tables]
  gl = hb.embedding.GroupLookup(dev_tables, ['mean'] * 26)
  opt = hb.training.Adagrad(lr)
  offsets = np.arange(B + 1, dtype=np.int64)
  tol = [None] * 26
  for step in range(2):
    ids = [bench.gen_ids_numpy(rng, B, n, "zipf", 1.05, salt=k) for k, n in enumerate(sizes)]
    grad = (rng.randn(B, 26 * D) * 1e-2).astype(np.float32)
    out = gl.forward([torch.from_numpy(i).cuda() for i in ids]).cpu().numpy()
    gl.backward_update(torch.from_numpy(grad).cuda(), opt, check=True)
    for k in range(26):
      exp = oracle.embedding_lookup_sparse(tables[k], ids[k], offsets, 'mean')
      np.testing.assert_allclose(out[:, k * D:(k + 1) * D], exp, rtol=1e-5, atol=1e-7 if step == 0 else 1e-6,
                                 err_msg=f'step {step} forward feature {k}')
      g = np.ascontiguousarray(grad[:, k * D:(k + 1) * D])
      tw, ta = rank_cases.adagrad_reference(oracle, tables[k], accs[k], ids[k], g, lr, tol[k])
      tol[k] = (tw, ta)
      got = dev_tables[k].cpu().numpy()
      bad = np.abs(got.astype(np.float64) - tables[k]) > tw
      assert not bad.any(), f'step {step} table {k} ({sizes[k]} rows): {int(bad.sum())} beyond the bound'



@pytest.mark.parametrize('dim', [4, 16, 64, 128, 256])
def test_adagrad_dims(hb, oracle, dim):
  _run(hb, oracle, 'adagrad', 'sum', [5000, 700], dim, 3000, steps=1)


def test_lazy_adam(hb, oracle):
  _run(hb, oracle, 'lazy_adam', 'mean', [20000, 300], 16, 5000, steps=3)
  _run(hb, oracle, 'lazy_adam', 'sum', [100000], 128, 4096, steps=2, one_hot=True, zipf=True, rtol=1e-4)


def test_determinism(hb):
  """Same inputs -> bit-identical tables (no float atomics)."""
  res = []
  for _ in range(2):
    g = torch.Generator(device='cuda').manual_seed(7)
    t = torch.randn(1000, 32, device='cuda', generator=g)
    ids = torch.randint(0, 1000, (50000,), device='cuda', generator=g)
    grad = torch.randn(50000, 32, device='cuda', generator=g)
    gl = hb.embedding.GroupLookup([t])
    gl.forward([ids])
    gl.backward_update(grad, hb.training.Adagrad(0.01))
    res.append(t.clone())
  assert torch.equal(res[0], res[1])


def test_untouched_rows_unchanged_and_oob(hb):
  t = torch.randn(100, 8, device='cuda')
  t0 = t.clone()
  gl = hb.embedding.GroupLookup([t], ['sum'])
  ids = torch.tensor([5, 7, 5], device='cuda')
  gl.forward([ids])
  gl.backward_update(torch.ones(3, 8, device='cuda'), hb.training.SGD(0.5), check=True)
  mask = torch.ones(100, dtype=torch.bool)
  mask[[5, 7]] = False
  assert torch.equal(t.cpu()[mask], t0.cpu()[mask])
  assert torch.allclose(t[5], t0[5] - 1.0) and torch.allclose(t[7], t0[7] - 0.5)


@pytest.mark.parametrize('rows', [511, 512, 513, 2 ** 18 - 1, 2 ** 18])
def test_invalid_id_does_not_split_the_run_of_a_valid_row(hb, rows):
  """Table sizes at a radix digit boundary: the sentinel of an out-of-range id must
  sort behind every valid row, not into the middle of the last row's run (ADVICE
  round 1)."""
  t = torch.zeros(rows, 8, device='cuda')
  gl = hb.embedding.GroupLookup([t], ['sum'])
  last = rows - 2
  ids = torch.tensor([last, -1, last, rows + 5, last, rows - 1], device='cuda')
  gl.forward([ids])
  gl.backward_update(torch.ones(6, 8, device='cuda'), hb.training.SGD(1.0))
  with pytest.raises(IndexError):
    hb._util.check_status(t.device)
  exp = torch.zeros(rows, 8)
  exp[last] = -3.0
  exp[rows - 1] = -1.0
  assert torch.equal(t.cpu(), exp)
