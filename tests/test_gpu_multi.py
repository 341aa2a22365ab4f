"""Multi-GPU parity (needs >= 2 GPUs; one process per GPU, bootstrap over gloo):
K2 AlltoallvN against the reference's golden vectors and the oracle, and the fused
sharded GroupLookup (forward and backward + Adagrad) against the unsharded oracle."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, fn_name, q):
  sys.path.insert(0, ROOT)
  try:
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('cpu:gloo,cuda:nccl', rank=rank, world_size=world)
    import hybridbackend_b200 as hb
    from oracle import hb_oracle as o
    globals()[fn_name](rank, world, hb, o)
    torch.cuda.synchronize()
    dist.barrier()
    q.put((rank, 'ok'))
  except Exception:  # pylint: disable=broad-except
    q.put((rank, traceback.format_exc()))
  finally:
    try:
      dist.destroy_process_group()
    except Exception:  # pylint: disable=broad-except
      pass


def _spawn(fn_name, world=2):
  if torch.cuda.device_count() < world:
    pytest.skip(f'needs {world} GPUs')
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
  for p in procs:
    p.start()
  import queue as _queue
  res = []
  try:
    for _ in range(world):
      try:
        res.append(q.get(timeout=150 if not res else 40))
      except _queue.Empty:
        res.append((-1, 'a rank never reported (hung in a kernel or crashed hard)'))
        break
  finally:
    for p in procs:
      p.join(timeout=10)
      if p.is_alive():
        p.kill()
  bad = [r for r in res if r[1] != 'ok']
  assert not bad, '\n'.join(f'rank {r}: {m}' for r, m in bad)


# ------------------------------------------------------------------------------
def _alltoallv_golden(rank, world, hb, o):
  coll = hb.distribute.Collective(rank, world, window_bytes=8 << 20)
  # alltoall_test.py:219-226
  ids = [[1, 2, 3], [4, 5, 6]]
  sizes = [[1, 2], [1, 2]]
  out, osz = coll.alltoall(torch.tensor(ids[rank], device='cuda'),
                           sizes=torch.tensor(sizes[rank], dtype=torch.int32, device='cuda'))
  exp_ids = [[1, 4], [2, 3, 5, 6]]
  exp_sz = [[1, 1], [2, 2]]
  assert out.tolist() == exp_ids[rank] and osz.tolist() == exp_sz[rank]
  # alltoall_test.py:254-269 (alltoallv_n, the enabled reference test)
  inputs = {0: [([1., 2., 3.], [1, 2]), ([4., 5., 6.], [2, 1])],
            1: [([7., 8., 9.], [2, 1]), ([10., 11., 12.], [1, 2])]}
  exp = {0: [([1., 7., 8.], [1, 2]), ([4., 5., 10.], [2, 1])],
         1: [([2., 3., 9.], [2, 1]), ([6., 11., 12.], [1, 2])]}
  vals = [torch.tensor(v, device='cuda') for v, _ in inputs[rank]]
  szs = [torch.tensor(s, dtype=torch.int32, device='cuda') for _, s in inputs[rank]]
  outs, oszs = coll.alltoall(vals, sizes=szs)
  for k in range(2):
    assert outs[k].tolist() == exp[rank][k][0] and oszs[k].tolist() == exp[rank][k][1]
  # alltoall_test.py:245-252 / :271-286: float payload over a float16 wire
  out, osz = coll.alltoall(torch.tensor([float(v) for v in ids[rank]], device='cuda'),
                           sizes=torch.tensor(sizes[rank], dtype=torch.int32, device='cuda'),
                           wire_dtype=torch.float16)
  assert out.dtype == torch.float32 and out.tolist() == [float(v) for v in exp_ids[rank]]
  assert osz.tolist() == exp_sz[rank]
  outs, oszs = coll.alltoall(vals, sizes=szs, wire_dtype=torch.float16)
  for k in range(2):
    assert outs[k].tolist() == exp[rank][k][0] and oszs[k].tolist() == exp[rank][k][1]
  # equal-split alltoall (alltoall_test.py:200-205): expected = transpose of inputs
  full = [torch.arange(6, dtype=torch.float32).reshape(2, 3) + 10 * d for d in range(world)]
  got = coll.alltoall(full[rank].cuda())
  assert torch.equal(got.cpu(), torch.stack([full[d][rank] for d in range(world)]))
  coll.barrier()
  coll.close()


def test_alltoallv_golden_2gpu():
  _spawn('_alltoallv_golden')


def _alltoallv_random(rank, world, hb, o):
  coll = hb.distribute.Collective(rank, world, window_bytes=256 << 20)
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
    vals = [torch.from_numpy(ins[k][rank]).cuda() for k in range(N)]
    szs = [torch.from_numpy(all_sizes[k, rank]).cuda() for k in range(N)]
    outs, oszs = coll.alltoall(vals, sizes=szs, common_shape=[dims[trial]] * N)
    hb._util.check_status(torch.device('cuda', rank))
    for k in range(N):
      eo, es = o.alltoallv(ins[k], all_sizes[k], dims[trial])
      assert np.array_equal(outs[k].cpu().numpy(), eo[rank]), (trial, k)
      assert np.array_equal(oszs[k].cpu().numpy(), es[rank])
  hb._util.check_status(torch.device('cuda', rank))
  coll.close()


def test_alltoallv_vs_oracle_2gpu():
  _spawn('_alltoallv_random')


class _Soft:
  """Collects assertion failures instead of raising mid-protocol: a rank that
  stops early would leave its peer spinning on a flag inside a kernel."""

  def __init__(self):
    self.errors = []

  def allclose(self, got, exp, msg, **kw):
    try:
      np.testing.assert_allclose(got, exp, err_msg=msg, **kw)
    except AssertionError as e:
      self.errors.append(str(e)[:600])

  def done(self):
    assert not self.errors, '\n'.join(self.errors)


def _sharded_lookup(rank, world, hb, o):
  dev = torch.device('cuda', rank)
  soft = _Soft()
  rng = np.random.RandomState(11)  # shared
  sizes = [1003, 40000, 2, 250000]   # 2 rows <= W -> "small" (replicated) table
  D, B = 32, 3000
  full = [rng.uniform(-0.1, 0.1, (n, D)).astype(np.float32) for n in sizes]
  feats_all = []
  for r in range(world):
    fr = []
    for j, n in enumerate(sizes):
      if j == 0:   # CSR bags
        lens = rng.poisson(2, B)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        fr.append(((rng.zipf(1.3, int(off[-1])) % n).astype(np.int64), off))
      else:
        fr.append((rng.randint(0, n, B).astype(np.int64), None))
    feats_all.append(fr)
  grads = [rng.randn(B, len(sizes) * D).astype(np.float32) for _ in range(world)]
  from hybridbackend_b200.embedding.sharded import plan_window_bytes
  max_nnz = [max(len(feats_all[r][j][0]) for r in range(world)) + 5 for j in range(len(sizes))]
  tables = [hb.embedding.ShardedEmbeddingWeights(f't{j}', n, D, rank, world, device=dev)
            for j, n in enumerate(sizes)]
  for t, f in zip(tables, full):
    t.load_global(torch.from_numpy(f))
  sh = [j for j, t in enumerate(tables) if t.sharded]
  assert sh == [0, 1, 3]
  wb = plan_window_bytes(world, [max_nnz[j] for j in sh], [D] * len(sh), world)
  coll = hb.distribute.Collective(rank, world, window_bytes=wb)
  comb = ['mean', 'sum', 'sqrtn', 'mean']
  gl = hb.embedding.GroupLookup(tables, comb, collective=coll, max_nnz=max_nnz, capacity_factor=world)
  opt = hb.training.Adagrad(0.05)
  ref_tables = [f.copy() for f in full]
  ref_acc = [np.full_like(f, 0.1) for f in full]
  for step in range(2):
    mine = feats_all[rank]
    out = gl.forward([torch.from_numpy(f[0]).to(dev) for f in mine],
                     [torch.from_numpy(f[1]).to(dev) if f[1] is not None else None for f in mine]
                     ).cpu().numpy()
    for j in range(len(sizes)):
      ids, off = mine[j]
      offs = off if off is not None else np.arange(len(ids) + 1, dtype=np.int64)
      exp = o.embedding_lookup_sparse(ref_tables[j], ids, offs, comb[j])
      # from the second step on the tables carry the (summation-order) differences of
      # the previous update, so the forward is compared with an absolute tolerance
      soft.allclose(out[:, j * D:(j + 1) * D], exp, f'step {step} feature {j}', rtol=1e-5,
                    atol=1e-7 if step == 0 else 2e-5)
    gl.backward_update(torch.from_numpy(grads[rank]).to(dev), opt)
    # oracle: sharded tables get the SUM over ranks of the per-rank gradients
    # (training/gradient.py:216-217), applied once per step and unique row
    for j in range(len(sizes)):
      rows, rgs = [], []
      for r in range(world):
        ids, off = feats_all[r][j]
        offs = off if off is not None else np.arange(len(ids) + 1, dtype=np.int64)
        g = np.ascontiguousarray(grads[r][:, j * D:(j + 1) * D])
        rows.append(ids)
        rgs.append(o.lookup_row_grads(g, offs, comb[j]))
      # sharded tables: owner-side order is by source rank then partitioned position;
      # replicated tables: all-gathered in rank order (gradient.py:163-177).  Either
      # way every row gets the sum over ALL ranks; association may differ.
      o.sparse_apply_adagrad(ref_tables[j], ref_acc[j], np.concatenate(rows), np.concatenate(rgs), 0.05)
    for j in sh:
      got = tables[j].weight.cpu().numpy()
      soft.allclose(got, ref_tables[j][rank::world], f'step {step} table {j}', rtol=2e-4, atol=1e-6)
    # small (replicated) table: gradients of all ranks all-gathered, replicas identical
    got = tables[2].weight.cpu().numpy()
    soft.allclose(got, ref_tables[2], f'step {step} replicated table', rtol=2e-4, atol=1e-6)
  torch.cuda.synchronize()
  dist.barrier()
  try:
    hb._util.check_status(dev)
  except Exception as e:  # pylint: disable=broad-except
    soft.errors.append(f'status word: {e}')
  coll.close()
  soft.done()


def test_sharded_group_lookup_2gpu():
  _spawn('_sharded_lookup')


def _sharded_overflow(rank, world, hb, o):
  dev = torch.device('cuda', rank)
  D, B, n = 16, 4096, 100000
  t = hb.embedding.ShardedEmbeddingWeights('t', n, D, rank, world, device=dev)
  t.weight.zero_()
  from hybridbackend_b200.embedding.sharded import plan_window_bytes
  coll = hb.distribute.Collective(rank, world, window_bytes=plan_window_bytes(world, [B], [D], 1.0))
  gl = hb.embedding.GroupLookup([t], ['sum'], collective=coll, max_nnz=[B], capacity_factor=1.0)
  ids = torch.zeros(B, dtype=torch.int64, device=dev)  # every id owned by rank 0: overflow
  gl.forward([ids])
  torch.cuda.synchronize()
  dist.barrier()
  if rank == 0:
    with pytest.raises(RuntimeError, match='overflow'):
      hb._util.check_status(dev)
  coll.close()


def test_sharded_overflow_is_reported_2gpu():
  _spawn('_sharded_overflow')
