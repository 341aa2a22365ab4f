"""Multi-rank parity cases, written once and run two ways:

  * tests/test_gpu_multi.py        one PROCESS per GPU (needs >= 2 GPUs), bootstrap over gloo;
  * tests/test_gpu_local_group.py  one THREAD per rank on ONE GPU through the in-process
                                   group (hbCommCreateLocalGroup): same entry points, same
                                   kernels, peer windows = each other's allocations.

A case is `fn(env)`; `env` gives rank/world/device, `collective(window_bytes)` (the
rank's communicator, created collectively) and `barrier()`.
"""
import numpy as np
import torch

U = 2.0 ** -24  # fp32 unit roundoff


class Soft:
  """Collects assertion failures instead of raising mid-protocol: a rank that
  stops early would leave its peers waiting for it."""

  def __init__(self):
    self.errors = []

  def check(self, cond, msg):
    if not cond:
      self.errors.append(msg)

  def allclose(self, got, exp, msg, **kw):
    try:
      np.testing.assert_allclose(got, exp, err_msg=msg, **kw)
    except AssertionError as e:
      self.errors.append(str(e)[:600])

  def within(self, got, exp, tol, msg):
    bad = np.abs(got.astype(np.float64) - exp.astype(np.float64)) > tol
    if bad.any():
      i = np.argwhere(bad)[0]
      self.errors.append(f'{msg}: {int(bad.sum())} elements beyond the bound, first at {tuple(i)}: '
                         f'got {got[tuple(i)]!r} expected {exp[tuple(i)]!r} tol {tol[tuple(i)]!r}')

  def done(self):
    assert not self.errors, '\n'.join(self.errors)


# ----------------------------------------------------------------------------------------
# fp32 summation-order bound of the fused update (the "hot row" tolerance)
# ----------------------------------------------------------------------------------------
def adagrad_reference(oracle, table, acc, rows_idx, row_grads, lr, prev=None):
  """Applies hbo_sparse_apply_adagrad in place and returns the per-element tolerances
  (tol_w, tol_acc) for an implementation that adds the duplicates of a row in a
  different (but fixed) order.  Per row with n gradient terms g_i this step:
    any two fp32 summation orders differ by  dG <= 2 (n-1) u sum|g_i|      (u = 2^-24)
    acc' = acc + G^2             ->  dAcc' <= dAcc + 2 |G| dG + dG^2 + 2 u acc'
    w'   = w - lr G / sqrt(acc') ->  dW'   <= dW + lr (dG / sqrt(acc') + |G| dAcc' / (2 acc'^1.5))
                                              + 4 u (|w'| + lr |G| / sqrt(acc'))
  where (dW, dAcc) = `prev` are the tolerances carried over from the previous step
  (None: the states start identical).  n = 1 gives a few ulp; n ~ 1e4 gives ~1e-3 of
  the step -- what the blanket rtol of round 1 hid.  Rows without a gradient keep
  their carried tolerance."""
  rows = table.shape[0]
  n = np.bincount(rows_idx, minlength=rows).astype(np.float64)
  S = np.zeros(table.shape, np.float64)
  np.add.at(S, rows_idx, np.abs(row_grads.astype(np.float64)))
  G = np.zeros(table.shape, np.float64)
  np.add.at(G, rows_idx, row_grads.astype(np.float64))
  oracle.sparse_apply_adagrad(table, acc, rows_idx, row_grads, lr)
  pw, pa = prev if prev is not None else (np.zeros(table.shape), np.zeros(table.shape))
  touched = (n > 0)[:, None]
  accn = acc.astype(np.float64)
  dG = 2.0 * np.maximum(n - 1, 0)[:, None] * U * S
  dAcc = pa + np.where(touched, 2.0 * np.abs(G) * dG + dG * dG + 2.0 * U * accn, 0.0)
  step = lr * np.abs(G) / np.sqrt(accn)
  dW = pw + np.where(touched, lr * (dG / np.sqrt(accn) + np.abs(G) * dAcc / (2.0 * accn ** 1.5)) +
                     4.0 * U * (np.abs(table) + step), 0.0)
  return dW + 1e-12, dAcc + 1e-12


# ----------------------------------------------------------------------------------------
# K2 AlltoallvN
# ----------------------------------------------------------------------------------------
def alltoallv_golden(env):
  """The reference's own vectors: alltoall_test.py:219-226 (alltoallv), :254-269
  (alltoallv_n), :245-252/:271-286 (fp16 wire), :200-205 (equal split)."""
  assert env.world == 2
  rank, dev = env.rank, env.device
  coll = env.collective(8 << 20)
  ids = [[1, 2, 3], [4, 5, 6]]
  sizes = [[1, 2], [1, 2]]
  out, osz = coll.alltoall(torch.tensor(ids[rank], device=dev),
                           sizes=torch.tensor(sizes[rank], dtype=torch.int32, device=dev))
  exp_ids = [[1, 4], [2, 3, 5, 6]]
  exp_sz = [[1, 1], [2, 2]]
  soft = Soft()
  soft.check(out.tolist() == exp_ids[rank] and osz.tolist() == exp_sz[rank], f'alltoallv: {out.tolist()} {osz.tolist()}')
  inputs = {0: [([1., 2., 3.], [1, 2]), ([4., 5., 6.], [2, 1])],
            1: [([7., 8., 9.], [2, 1]), ([10., 11., 12.], [1, 2])]}
  exp = {0: [([1., 7., 8.], [1, 2]), ([4., 5., 10.], [2, 1])],
         1: [([2., 3., 9.], [2, 1]), ([6., 11., 12.], [1, 2])]}
  vals = [torch.tensor(v, device=dev) for v, _ in inputs[rank]]
  szs = [torch.tensor(s, dtype=torch.int32, device=dev) for _, s in inputs[rank]]
  outs, oszs = coll.alltoall(vals, sizes=szs)
  for k in range(2):
    soft.check(outs[k].tolist() == exp[rank][k][0] and oszs[k].tolist() == exp[rank][k][1], f'alltoallv_n {k}')
  out, osz = coll.alltoall(torch.tensor([float(v) for v in ids[rank]], device=dev),
                           sizes=torch.tensor(sizes[rank], dtype=torch.int32, device=dev),
                           wire_dtype=torch.float16)
  soft.check(out.dtype == torch.float32 and out.tolist() == [float(v) for v in exp_ids[rank]], 'fp16 wire')
  soft.check(osz.tolist() == exp_sz[rank], 'fp16 wire sizes')
  outs, oszs = coll.alltoall(vals, sizes=szs, wire_dtype=torch.float16)
  for k in range(2):
    soft.check(outs[k].tolist() == exp[rank][k][0] and oszs[k].tolist() == exp[rank][k][1], f'fp16 n {k}')
  full = [torch.arange(6, dtype=torch.float32).reshape(2, 3) + 10 * d for d in range(env.world)]
  got = coll.alltoall(full[rank].to(dev))
  soft.check(torch.equal(got.cpu(), torch.stack([full[d][rank] for d in range(env.world)])), 'equal split')
  coll.barrier()
  torch.cuda.synchronize()
  env.barrier()
  coll.close()
  soft.done()


def alltoallv_grads(env):
  """Gradient vectors of the reference: alltoall_test.py:207-217 (alltoall),
  :228-243 (alltoallv), :288-304 (alltoallv_n): loss = mean(output), upstream g.
  With one process per rank the gradient flows through torch autograd
  (_AlltoallvFn); with one THREAD per rank the backward is called directly
  (Collective.alltoall_grad, the call autograd makes) because torch runs all CUDA
  backward nodes of a device on one engine thread, where two ranks' collectives
  cannot meet."""
  assert env.world == 2
  rank, dev = env.rank, env.device
  coll = env.collective(8 << 20)
  soft = Soft()
  g = 2.0

  def grad_of(x, sizes):
    if env.autograd:
      x = x.clone().requires_grad_(True)
      out = coll.alltoall(x, sizes=sizes)
      out = out[0] if sizes is not None else out
      (out.mean() * g).backward()
      return x.grad
    out = coll.alltoall(x, sizes=sizes)
    out, osz = out if sizes is not None else (out, None)
    dout = torch.full_like(out, g / out.numel())
    return coll.alltoall_grad(dout, osz)

  # alltoallv_grad: sizes [[5,1],[3,4]]
  sizes = [[5, 1], [3, 4]]
  values = [3.6, 4.2]
  x = torch.full((sum(sizes[rank]),), values[rank], device=dev)
  got = grad_of(x, torch.tensor(sizes[rank], dtype=torch.int32, device=dev))
  g0 = g / (sizes[0][0] + sizes[1][0])
  g1 = g / (sizes[0][1] + sizes[1][1])
  exp = sizes[rank][0] * [g0] + sizes[rank][1] * [g1]
  soft.allclose(got.cpu().numpy(), np.asarray(exp, np.float32), 'alltoallv grad', rtol=1e-6)
  # alltoallv_n_grad: every gradient is g / 3
  inputs = {0: [([1., 2., 3.], [1, 2]), ([4., 5., 6.], [2, 1])],
            1: [([7., 8., 9.], [2, 1]), ([10., 11., 12.], [1, 2])]}
  for v, s in inputs[rank]:
    got = grad_of(torch.tensor(v, device=dev), torch.tensor(s, dtype=torch.int32, device=dev))
    soft.allclose(got.cpu().numpy(), np.full(3, g / 3, np.float32), 'alltoallv_n grad', rtol=1e-6)
  # alltoall_grad (equal split, [w, h] = [2, 10]): g / (w * h) everywhere
  got = grad_of(torch.randn(2, 10, device=dev), None)
  soft.allclose(got.cpu().numpy(), np.full((2, 10), g / 20, np.float32), 'alltoall grad', rtol=1e-6)
  torch.cuda.synchronize()
  env.barrier()
  coll.close()
  soft.done()


def alltoallv_random(env):
  rank, world, dev, o = env.rank, env.world, env.device, env.oracle
  coll = env.collective(256 << 20)
  soft = Soft()
  rng = np.random.RandomState(0)  # same stream on every rank
  for trial in range(4):
    N = [1, 3, 26, 5][trial]
    dims = [(), (16,), (64,), (3, 5)]
    dts = [np.int64, np.float32, np.float32, np.int32]
    all_sizes = rng.randint(0, [5, 3000, 4000, 40][trial], size=(N, world, world)).astype(np.int32)
    if trial == 1:
      all_sizes[0, :, :] = 0  # an all-empty tensor
    ins = [[rng.randint(-1000, 1000, size=(int(all_sizes[k, r].sum()),) + dims[trial]).astype(dts[trial])
            for r in range(world)] for k in range(N)]
    vals = [torch.from_numpy(ins[k][rank]).to(dev) for k in range(N)]
    szs = [torch.from_numpy(all_sizes[k, rank]).to(dev) for k in range(N)]
    outs, oszs = coll.alltoall(vals, sizes=szs, common_shape=[dims[trial]] * N)
    for k in range(N):
      eo, es = o.alltoallv(ins[k], all_sizes[k], dims[trial])
      soft.check(np.array_equal(outs[k].cpu().numpy(), eo[rank]), f'trial {trial} tensor {k} payload')
      soft.check(np.array_equal(oszs[k].cpu().numpy(), es[rank]), f'trial {trial} tensor {k} sizes')
  torch.cuda.synchronize()
  env.barrier()
  coll.close()
  soft.done()


def alltoallv_overflow(env):
  """A receive volume beyond half of the window raises on EVERY rank instead of
  returning uninitialised outputs."""
  rank, world, dev = env.rank, env.world, env.device
  coll = env.collective(1 << 20)
  n = 200000  # 800 KB per segment > 512 KB half window
  x = torch.zeros(n * world, dtype=torch.float32, device=dev)
  sizes = torch.full((world,), n, dtype=torch.int32, device=dev)
  err = None
  try:
    coll.alltoall(x, sizes=sizes)
  except RuntimeError as e:
    err = str(e)
  torch.cuda.synchronize()
  env.barrier()
  coll.close()
  assert err is not None and 'window' in err, err


def allreduce_case(env):
  rank, world, dev = env.rank, env.world, env.device
  coll = env.collective(16 << 20)
  soft = Soft()
  rng = np.random.RandomState(5)
  for count in [1, 7, 1000, 100003]:
    xs = [rng.randn(count).astype(np.float32) for _ in range(world)]
    got = coll.allreduce(torch.from_numpy(xs[rank]).to(dev), scale=1.0 / world)
    exp = xs[0].copy()
    for q in range(1, world):
      exp = exp + xs[q]                  # rank order, fp32
    exp = exp * np.float32(1.0 / world)
    soft.check(np.array_equal(got.cpu().numpy(), exp), f'allreduce count {count}')
  torch.cuda.synchronize()
  env.barrier()
  coll.close()
  soft.done()


# ----------------------------------------------------------------------------------------
# fused sharded GroupLookup: forward + backward + Adagrad vs the unsharded oracle
# ----------------------------------------------------------------------------------------
def _sharded_train(env, sizes, D, B, comb, gen_feature, steps=2, lr=0.05, capacity_factor=None,
                   batch_size=-1, seed=11):
  """Every rank looks up its own batch in tables sharded over the ranks, then applies
  Adagrad.  Oracle: the UNSHARDED computation -- embedding_lookup_sparse per rank on the
  full tables; sharded tables receive the SUM over ranks of the row gradients
  (training/gradient.py:216-217), replicated small tables the dense MEAN (:132-141,
  :157-160, :77-97)."""
  rank, world, dev, hb, o = env.rank, env.world, env.device, env.hb, env.oracle
  soft = Soft()
  rng = np.random.RandomState(seed)  # shared stream: every rank generates everything
  F = len(sizes)
  full = [rng.uniform(-0.1, 0.1, (n, D)).astype(np.float32) for n in sizes]
  from hybridbackend_b200.embedding.sharded import plan_window_bytes
  tables = [hb.embedding.ShardedEmbeddingWeights(f't{j}', n, D, rank, world, batch_size=batch_size, device=dev)
            for j, n in enumerate(sizes)]
  for t, f in zip(tables, full):
    t.load_global(torch.from_numpy(f))
  sh = [j for j, t in enumerate(tables) if t.sharded]
  rep = [j for j in range(F) if j not in sh]
  ref_tables = [f.copy() for f in full]
  ref_acc = [np.full_like(f, 0.1) for f in full]
  gl = None
  coll = None
  opt = hb.training.Adagrad(lr)
  for step in range(steps):
    feats_all = [[gen_feature(rng, j, n, B) for j, n in enumerate(sizes)] for _ in range(world)]
    grads = [rng.randn(B, F * D).astype(np.float32) for _ in range(world)]
    if gl is None:
      max_nnz = [max(len(feats_all[r][j][0]) for r in range(world)) * 2 + 8 for j in range(F)]
      cf = capacity_factor or world
      wb = plan_window_bytes(world, [max_nnz[j] for j in sh], [D] * len(sh), cf) + (32 << 20)
      coll = env.collective(wb)
      gl = hb.embedding.GroupLookup(tables, comb, collective=coll, max_nnz=max_nnz, capacity_factor=cf)
    mine = feats_all[rank]
    out = gl.forward([torch.from_numpy(f[0]).to(dev) for f in mine],
                     [torch.from_numpy(f[1]).to(dev) if f[1] is not None else None for f in mine]
                     ).cpu().numpy()
    for j in range(F):
      ids, off = mine[j]
      offs = off if off is not None else np.arange(len(ids) + 1, dtype=np.int64)
      exp = o.embedding_lookup_sparse(ref_tables[j], ids, offs, comb[j])
      # forward: same rows, same bag order -> the pooled rows equal the oracle's up to
      # the table differences the previous update left (bounded below)
      soft.allclose(out[:, j * D:(j + 1) * D], exp, f'step {step} feature {j} forward', rtol=1e-5,
                    atol=1e-7 if step == 0 else 5e-5)
    gl.backward_update(torch.from_numpy(grads[rank]).to(dev), opt)
    tol = dict(prev_tol) if step else {}
    for j in range(F):
      rows, rgs = [], []
      for r in range(world):
        ids, off = feats_all[r][j]
        offs = off if off is not None else np.arange(len(ids) + 1, dtype=np.int64)
        g = np.ascontiguousarray(grads[r][:, j * D:(j + 1) * D])
        rows.append(ids)
        rgs.append(o.lookup_row_grads(g, offs, comb[j]))
      if j in sh:
        tol[j] = adagrad_reference(o, ref_tables[j], ref_acc[j], np.concatenate(rows),
                                   np.concatenate(rgs), lr, tol.get(j))
      else:
        # dense gradient per rank (position order), rank-order sum, * 1/W, dense apply
        n = sizes[j]
        dense = np.zeros((n, D), np.float32)
        S = np.zeros((n, D), np.float64)
        cnt = np.zeros(n, np.float64)
        for r in range(world):
          d = np.zeros((n, D), np.float32)
          np.add.at(d, rows[r], rgs[r])  # position order within the rank
          dense = dense + d if r else d
          np.add.at(S, rows[r], np.abs(rgs[r].astype(np.float64)))
          cnt += np.bincount(rows[r], minlength=n)
        dense = dense * np.float32(1.0 / world)
        # bound: treat the mean gradient as one term whose own error is the sum bound / W
        tw, ta = adagrad_reference(o, ref_tables[j], ref_acc[j], np.arange(n, dtype=np.int64), dense, lr,
                                   tol.get(j))
        dG = 2.0 * np.maximum(cnt - 1, 0)[:, None] * U * S / world
        accn = ref_acc[j].astype(np.float64)
        ta = ta + 2.0 * np.abs(dense) * dG + dG * dG
        tw = tw + 2.0 * lr * dG / np.sqrt(accn)
        tol[j] = (tw, ta)
    for j in sh:
      got = tables[j].weight.cpu().numpy()
      soft.within(got, ref_tables[j][rank::world], tol[j][0][rank::world], f'step {step} table {j}')
      soft.within(tables[j].slots[0].cpu().numpy(), ref_acc[j][rank::world], tol[j][1][rank::world],
                  f'step {step} accumulator {j}')
    for j in rep:
      soft.within(tables[j].weight.cpu().numpy(), ref_tables[j], tol[j][0], f'step {step} replicated table {j}')
    prev_tol = tol
  torch.cuda.synchronize()
  env.barrier()
  try:
    hb._util.check_status(dev)
  except Exception as e:  # pylint: disable=broad-except
    soft.errors.append(f'status word: {e}')
  gl.close()
  coll.close()
  soft.done()


def _gen_mixed(rng, j, n, B):
  if j == 0:   # CSR bags, zipf ids
    lens = rng.poisson(2, B)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return ((rng.zipf(1.3, int(off[-1])) % n).astype(np.int64), off)
  return (rng.randint(0, n, B).astype(np.int64), None)


def sharded_lookup(env):
  # 2 rows <= W -> "small" (replicated) table; CSR + one-hot features; mean/sum/sqrtn
  _sharded_train(env, [1003, 40000, 2, 250000], 32, 3000, ['mean', 'sum', 'sqrtn', 'mean'], _gen_mixed)


def sharded_lookup_dim64_hot(env):
  """C3's regime at test size: D=64, Criteo-like table sizes incl. tiny hot tables
  (sharded when rows > W), Zipf(1.05) ids as bench.py draws them."""
  import bench
  sizes = [39884406 // 400, 39043, 17289, 3, 7120, 63, 2953546 // 40, 10, 155, 4, 36, 976]

  def gen(rng, j, n, B):
    return (bench.gen_ids_numpy(rng, B, n, 'zipf', 1.05, salt=j), None)
  _sharded_train(env, sizes, 64, 4096, ['mean'] * len(sizes), gen, steps=2, lr=0.01)


def sharded_many_features(env):
  """C4's shape at test size: 200 features (beyond round 1's 64-per-plan limit), D=16,
  multi-id bags (mean length 3), Zipf(1.2) ids, log-uniform vocabularies."""
  rs = np.random.RandomState(3)
  sizes = [int(v) for v in np.exp(rs.uniform(np.log(50), np.log(200000), 200))]

  def gen(rng, j, n, B):
    lens = rng.randint(1, 6, B)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return ((rng.zipf(1.2, int(off[-1])) % n).astype(np.int64), off)
  _sharded_train(env, sizes, 16, 512, ['mean'] * 200, gen, steps=1, lr=0.01)


def sharded_hot_keys_lazy_adam(env):
  """C5's regime at test size (BASELINE configs[4]): 8 features, D=128, 90 % of the ids
  drawn from a fixed 1 % hot set, backward scatter-add + sparse (Lazy) Adam through the
  sharded path.  C5 as stated (1e9 rows x D128 fp32 + two Adam slots = 1.5 TB) exceeds
  8 x 180 GB; the row count is cut, the hot-key structure is what is tested.
  Oracle: unsharded LazyAdam on the concatenation of all ranks' row gradients; the
  requester-side pre-sums change the fp32 association only (rtol 1e-4 on m / v, the
  weight step is +-lr_t * m / (sqrt(v) + eps), insensitive to it: atol 2e-6)."""
  rank, world, dev, hb, o = env.rank, env.world, env.device, env.hb, env.oracle
  soft = Soft()
  rng = np.random.RandomState(5)
  F, D, B, n = 8, 128, 1024, 50000
  hot = rng.choice(n, n // 100, replace=False)
  full = [rng.uniform(-0.05, 0.05, (n, D)).astype(np.float32) for _ in range(F)]
  from hybridbackend_b200.embedding.sharded import plan_window_bytes
  tables = [hb.embedding.ShardedEmbeddingWeights(f'h{j}', n, D, rank, world, device=dev) for j in range(F)]
  for t, f in zip(tables, full):
    t.load_global(torch.from_numpy(f))
  coll = env.collective(plan_window_bytes(world, [B] * F, [D] * F, world) + (8 << 20))
  gl = hb.embedding.GroupLookup(tables, ['sum'] * F, collective=coll, max_nnz=[B] * F)
  opt = hb.training.LazyAdam(0.001)
  ref = [f.copy() for f in full]
  ref_m = [np.zeros_like(f) for f in full]
  ref_v = [np.zeros_like(f) for f in full]
  for step in range(2):
    ids_all = []
    for r in range(world):
      per = []
      for j in range(F):
        pick_hot = rng.random_sample(B) < 0.9
        per.append(np.where(pick_hot, hot[rng.randint(0, len(hot), B)], rng.randint(0, n, B)).astype(np.int64))
      ids_all.append(per)
    grads = [(rng.randn(B, F * D) * 1e-2).astype(np.float32) for _ in range(world)]
    out = gl.forward([torch.from_numpy(i).to(dev) for i in ids_all[rank]]).cpu().numpy()
    for j in range(F):
      soft.allclose(out[:, j * D:(j + 1) * D], ref[j][ids_all[rank][j]], f'step {step} feature {j} forward',
                    rtol=1e-5, atol=0 if step == 0 else 1e-5)
    gl.backward_update(torch.from_numpy(grads[rank]).to(dev), opt)
    for j in range(F):
      rows = np.concatenate([ids_all[r][j] for r in range(world)])
      rg = np.concatenate([grads[r][:, j * D:(j + 1) * D] for r in range(world)])
      o.sparse_apply_lazy_adam(ref[j], ref_m[j], ref_v[j], rows, rg, 0.001, 0.9, 0.999, 1e-8, step + 1)
      soft.allclose(tables[j].weight.cpu().numpy(), ref[j][rank::world], f'step {step} table {j}', rtol=0, atol=2e-6 * (step + 1))
      soft.allclose(tables[j].slots[0].cpu().numpy(), ref_m[j][rank::world], f'step {step} m {j}', rtol=1e-4, atol=1e-8)
      soft.allclose(tables[j].slots[1].cpu().numpy(), ref_v[j][rank::world], f'step {step} v {j}', rtol=2e-4, atol=1e-11)
  torch.cuda.synchronize()
  env.barrier()
  try:
    hb._util.check_status(dev)
  except Exception as e:  # pylint: disable=broad-except
    soft.errors.append(f'status word: {e}')
  gl.close()
  coll.close()
  soft.done()


def sharded_overflow(env):
  """Receive capacity exceeded at ONE owner: every rank's status word reports it."""
  rank, world, dev, hb = env.rank, env.world, env.device, env.hb
  D, B, n = 16, 4096, 100000
  t = hb.embedding.ShardedEmbeddingWeights('t', n, D, rank, world, device=dev)
  t.weight.zero_()
  from hybridbackend_b200.embedding.sharded import plan_window_bytes
  coll = env.collective(plan_window_bytes(world, [B], [D], 1.0) + (1 << 20))
  gl = hb.embedding.GroupLookup([t], ['sum'], collective=coll, max_nnz=[B], capacity_factor=1.0)
  # every rank asks owner 0 for B distinct rows: W*B > capacity B
  ids = (torch.arange(B, dtype=torch.int64, device=dev) * world)
  gl.forward([ids])
  torch.cuda.synchronize()
  env.barrier()
  err = None
  try:
    hb._util.check_status(dev)
  except RuntimeError as e:
    err = str(e)
  gl.close()
  coll.close()
  assert err is not None and 'overflow' in err, f'rank {rank}: {err}'


def sharded_plan_recreate(env):
  """A second plan on the same communicator starts with fresh epochs that the flags
  of the first plan cannot satisfy: results stay correct."""
  rank, world, dev, hb, o = env.rank, env.world, env.device, env.hb, env.oracle
  from hybridbackend_b200.embedding.sharded import plan_window_bytes
  rng = np.random.RandomState(21)
  n, D, B = 5000, 8, 700
  full = rng.randn(n, D).astype(np.float32)
  coll = env.collective(plan_window_bytes(world, [B], [D], world) + (1 << 20))
  soft = Soft()
  for round_ in range(3):
    t = hb.embedding.ShardedEmbeddingWeights('t', n, D, rank, world, device=dev)
    t.load_global(torch.from_numpy(full))
    gl = hb.embedding.GroupLookup([t], ['sum'], collective=coll, max_nnz=[B])
    for step in range(round_ + 1):  # different step counts per plan: epochs would collide
      ids_all = [rng.randint(0, n, B).astype(np.int64) for _ in range(world)]
      out = gl.forward([torch.from_numpy(ids_all[rank]).to(dev)]).cpu().numpy()
      soft.check(np.array_equal(out, full[ids_all[rank]]), f'plan {round_} step {step}')
    torch.cuda.synchronize()
    env.barrier()
    gl.close()
  coll.close()
  soft.done()


def bootstrap_from_unique_id(env):
  """The reference's bootstrap protocol: rank 0 makes ONE 128-byte id (HbGetNcclId), it is
  broadcast (here over torch.distributed, in the reference over TF), every rank creates
  its communicator from it (HbCreateNcclCollective) -- then an alltoallv runs over it."""
  import torch.distributed as dist
  rank, world, dev, hb = env.rank, env.world, env.device, env.hb
  box = [hb.distribute.Collective.get_unique_id() if rank == 0 else None]
  dist.broadcast_object_list(box, src=0)
  coll = hb.distribute.Collective(rank, world, window_bytes=8 << 20, device=dev, unique_id=box[0])
  sizes = torch.tensor([rank + 1 + q for q in range(world)], dtype=torch.int32, device=dev)
  x = torch.full((int(sizes.sum()),), float(rank), device=dev)
  out, osz = coll.alltoall(x, sizes)
  torch.cuda.synchronize()
  exp_sz = [q + 1 + rank for q in range(world)]
  exp = np.concatenate([np.full(exp_sz[q], float(q), np.float32) for q in range(world)])
  ok = np.array_equal(osz.cpu().numpy(), np.array(exp_sz, np.int32)) and np.array_equal(out.cpu().numpy(), exp)
  env.barrier()
  coll.close()
  assert ok, f'rank {rank}: alltoallv over the id-bootstrapped communicator is wrong'
