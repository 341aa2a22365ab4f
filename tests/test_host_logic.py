"""Host-side logic on CPU: sharding rule vs the oracle, optimizer descriptors,
and the world_size-2 recipe over gloo (the oracle's per-rank functions as the
compute, torch.distributed/gloo as the transport) against the in-process
W-rank simulator and the unsharded lookup."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_rule_matches_oracle(hb, oracle):
  from hybridbackend_b200.embedding import sharding
  for n in (1, 5, 8, 9, 65536, 65537, 39884406):
    for w in (1, 2, 3, 8):
      for s in range(w):
        assert sharding.shard_rows(n, w, s) == oracle.shard_rows(n, w, s)
        assert sharding.shard_offset(n, w, s) == oracle.shard_offset(n, w, s)
      for b in (-1, 100, 65536):
        assert sharding.is_small_table(n, w, b) == oracle.is_small_table(n, w, b)


def test_criteo_small_table_count(hb):
  """with batch_size=65536, 18 of the 26 Criteo tables replicate."""
  from conftest import criteo_table_sizes
  from hybridbackend_b200.embedding import sharding
  sizes = criteo_table_sizes()
  assert len(sizes) == 26
  assert sum(sharding.is_small_table(n, 8, 65536) for n in sizes) == 18
  assert sum(sharding.is_small_table(n, 8, -1) for n in sizes) == 2  # rows <= W only


def test_optimizer_descriptors(hb):
  a = hb.training.Adagrad(0.01)
  assert a.num_slots == 1 and a.slot_init(0) == pytest.approx(0.1)
  d = a.descriptor()
  assert d.kind == 1 and d.lr == pytest.approx(0.01)
  l = hb.training.LazyAdam(0.001)
  l.step = 3
  d = l.descriptor()
  assert d.kind == 2 and d.step == 3 and d.beta2 == pytest.approx(0.999)


def test_segment_ids_to_offsets(hb):
  seg = torch.tensor([0, 0, 2, 2, 2, 5])
  off = hb.embedding.segment_ids_to_offsets(seg, 7)
  assert off.tolist() == [0, 2, 2, 5, 5, 5, 6, 6]


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _gloo_worker(rank, world, port, q):
  import sys
  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  from oracle import hb_oracle as o
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    rng = np.random.RandomState(100)
    N, D = 1001, 8
    table = rng.randn(N, D).astype(np.float32)          # same on all ranks
    shard = np.ascontiguousarray(table[rank::world])    # embedding/variables.py:107-111
    ids = np.random.RandomState(200 + rank).randint(0, N, 300 + 17 * rank).astype(np.int64)
    # sharding.py:179-180 partition
    part, sizes, idx = o.partition_by_modulo(ids, world)
    # :181-182 alltoallv(ids): sizes exchange then payload, over gloo
    send_sizes = torch.from_numpy(sizes.astype(np.int64))
    recv_sizes = torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(recv_sizes, send_sizes)
    recv = torch.empty(int(recv_sizes.sum()), dtype=torch.int64)
    dist.all_to_all_single(recv, torch.from_numpy(part), recv_sizes.tolist(), send_sizes.tolist())
    # :183-192 owner side
    u, inv = o.unique(recv.numpy())
    emb = shard[u // world][inv]
    # :193-196 return alltoallv with the received sizes
    back = torch.empty(len(ids), D)
    dist.all_to_all_single(back, torch.from_numpy(np.ascontiguousarray(emb)),
                           send_sizes.tolist(), recv_sizes.tolist())
    out = back.numpy()[idx]                              # :197-199 stitch
    ok = bool(np.array_equal(out, table[ids]))
    # token all-gather used by Collective's bootstrap
    toks = [None] * world
    dist.all_gather_object(toks, bytes([rank]) * 128)
    ok = ok and b''.join(toks) == bytes([0]) * 128 + bytes([1]) * 128
    q.put((rank, ok, recv_sizes.tolist()))
  finally:
    dist.destroy_process_group()


def test_sharded_recipe_over_gloo_world2(oracle):
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in range(world)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert all(ok for _, ok, _ in res)
  # cross-check the exchanged sizes against the in-process simulator
  ids = [np.random.RandomState(200 + r).randint(0, 1001, 300 + 17 * r).astype(np.int64)
         for r in range(world)]
  sizes = np.stack([oracle.partition_by_modulo(i, world)[1] for i in ids])
  for rank, _, rs in res:
    assert rs == sizes[:, rank].tolist()


def test_checkpoint_layout_roundtrip(hb):
  """Merged checkpoint layout (variables.py:118-132): contiguous SaveSliceInfo
  blocks of row-interleaved shards; round trip and logical permutation."""
  from hybridbackend_b200.embedding import checkpoint, sharding
  for n, w in ((10, 4), (1003, 8), (7, 2)):
    table = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)   # logical order
    parts = [table[s::w] for s in range(w)]                           # sharding.py:185-186
    merged = checkpoint.merge_shards(parts, n)
    back = checkpoint.split_merged(merged, w)
    for s in range(w):
      assert torch.equal(back[s], parts[s])
      off = sharding.shard_offset(n, w, s)
      # merged row off + r holds logical id r*w + s
      for r in (0, parts[s].shape[0] - 1):
        assert torch.equal(merged[off + r], table[r * w + s])
    perm = checkpoint.logical_rows_of_merged(n, w)
    assert sorted(perm.tolist()) == list(range(n))
    logical = torch.empty_like(merged)
    logical[perm] = merged
    assert torch.equal(logical, table)


def test_bench_id_scramble_spreads_the_hot_rows_over_the_owners():
  """bench.gen_ids_numpy: with the per-feature salt the hottest id of each feature maps to a
  different row, so under id % W sharding no rank owns the hot row of every table (the
  round-1 scramble sent rank 1 of EVERY table to id 0 -> owner 0)."""
  import collections
  import bench
  W = 8
  owners = collections.Counter()
  for k, vocab in enumerate(bench.CRITEO_SIZES):
    if vocab <= W:
      continue
    ids = bench.gen_ids_numpy(np.random.RandomState(k), 20000, vocab, 'zipf', 1.05, salt=k)
    assert ids.min() >= 0 and ids.max() < vocab
    vals, cnt = np.unique(ids, return_counts=True)
    owners[int(vals[np.argmax(cnt)]) % W] += 1
  assert len(owners) >= 5 and max(owners.values()) <= 8, owners
  # salt 0 reproduces the unsalted stream (old fixtures stay valid)
  a = bench.gen_ids_numpy(np.random.RandomState(1), 1000, 12345, 'zipf', 1.05)
  b = bench.gen_ids_numpy(np.random.RandomState(1), 1000, 12345, 'zipf', 1.05, salt=0)
  assert np.array_equal(a, b)


def test_group_window_bytes_covers_plan_and_replicated_allreduce():
  """The Collective window a GroupLookup needs = sharded plan + both halves of the dense
  all-reduce of the replicated small tables (hbAllreduceSumF32: world x count floats each)."""
  from hybridbackend_b200.embedding.sharded import group_window_bytes, plan_window_bytes
  import bench
  W, B, D = 8, 65536, 64
  sizes = bench.CRITEO_SIZES
  sh = [n for n in sizes if n > W]
  plan = plan_window_bytes(W, [B] * len(sh), [D] * len(sh), W)
  total = group_window_bytes(W, sizes, [D] * len(sizes), [B] * len(sizes), W)
  dense_floats = sum(n * D for n in sizes if n <= W)
  assert dense_floats == 7 * D                       # the 3- and 4-row Criteo tables
  assert total >= plan + 2 * W * dense_floats * 4    # what failed at N=8 before the helper existed
  assert total - plan < (4 << 20)
  # one rank: nothing is sharded, nothing is all-reduced
  assert group_window_bytes(1, sizes, [D] * len(sizes), [B] * len(sizes)) <= (2 << 20)


def test_checkpoint_save_merge_restore_protocol(tmp_path):
  """training/saver.py:89-180 in a neutral container: every rank saves `<name>/part_<rank>`
  (+ slots) with its SaveSliceInfo, the chief merges, each rank restores its slice; the
  merged variable has the reference's layout (shard s at rows [shard_offset(s), +rows(s)),
  holding logical ids r * W + s)."""
  from hybridbackend_b200.embedding import checkpoint as ck
  from hybridbackend_b200.embedding.sharding import shard_rows
  W, n, D = 3, 100, 4
  logical = torch.arange(n * D, dtype=torch.float32).reshape(n, D)      # row g = embedding of id g
  acc = logical * 0.5
  prefix = str(tmp_path / 'model.ckpt-7')
  for r in range(W):
    part, apart = logical[r::W].contiguous(), acc[r::W].contiguous()
    assert part.shape[0] == shard_rows(n, W, r)
    p = ck.save_local_shards(prefix, 'abc123', r, W, {'emb': (n, part)}, {'emb': {'Adagrad': apart}})
    assert p.endswith(f'part-{r:05d}.npz')
  out = ck.merge_checkpoint(prefix, 'abc123', W)
  assert out == prefix + '.npz' and not (tmp_path / 'model.ckpt-7_temp_abc123').exists()
  with np.load(out) as z:
    merged = torch.from_numpy(z['emb'])
  assert torch.equal(merged, ck.merge_shards([logical[r::W] for r in range(W)], n))
  perm = ck.logical_rows_of_merged(n, W)
  assert torch.equal(merged, logical[perm])            # merged row m holds logical id perm[m]
  for r in range(W):
    got = ck.restore_local_shards(prefix, r, W)
    assert torch.equal(got['emb'], logical[r::W]) and torch.equal(got['emb/Adagrad'], acc[r::W])
  with pytest.raises(ValueError):
    ck.restore_local_shards(prefix, 0, W + 1)
  with pytest.raises(FileNotFoundError):
    ck.merge_checkpoint(prefix, 'missing', W)
