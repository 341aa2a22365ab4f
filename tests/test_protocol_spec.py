"""The offset algebra of the fused sharded exchange (sharded.cu's ShFeatMeta),
restated in numpy (oracle/protocol.py), against the reference recipe."""
import numpy as np
import pytest


@pytest.mark.parametrize('W', [1, 2, 3, 8])
def test_fused_protocol_matches_reference_recipe(oracle, W):
  from oracle import protocol
  rng = np.random.RandomState(W)
  N, D = 1009, 4
  table = rng.randn(N, D).astype(np.float32)
  shards = oracle.shard_table(table, W)
  ids = [(rng.zipf(1.3, rng.randint(0, 400)) % N).astype(np.int64) for _ in range(W)]
  rows, ids_in, meta = protocol.fused_forward(shards, N, ids)
  ref = oracle.sharded_embedding_lookup(shards, N, ids)
  for r in range(W):
    assert np.array_equal(rows[r], ref[r])
    assert np.array_equal(rows[r], table[ids[r]])
    # every id an owner received is one it owns (id % W == r), rank-major order
    assert np.all(ids_in[r] % W == r)
    # received stream == oracle alltoallv of the partitioned ids
  parts = [oracle.partition_by_modulo(i, W) for i in ids]
  recv, _ = oracle.alltoallv([p[0] for p in parts], [p[1] for p in parts])
  for r in range(W):
    assert np.array_equal(ids_in[r], recv[r])


@pytest.mark.parametrize('W', [2, 8])
def test_dedup_variant_same_rows_fewer_wire_ids(oracle, W):
  from oracle import protocol
  rng = np.random.RandomState(10 + W)
  N, D = 5000, 4
  table = rng.randn(N, D).astype(np.float32)
  shards = oracle.shard_table(table, W)
  ids = [(rng.zipf(1.2, 3000) % N).astype(np.int64) for _ in range(W)]
  rows, wire = protocol.fused_forward_dedup(shards, N, ids)
  for r in range(W):
    assert np.array_equal(rows[r], table[ids[r]])
  assert wire < sum(len(i) for i in ids) / 2      # zipf: most ids are duplicates
