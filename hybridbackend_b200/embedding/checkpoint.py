"""Checkpoint layout of sharded embedding tables.

The reference saves shard s under `<name>/part_<s>` with a SaveSliceInfo that
claims the CONTIGUOUS row range [shard_offset(s), +shard_rows(s)) of the full
`[bucket_size, dim]` variable (hybridbackend/tensorflow/embedding/variables.py:
118-132), although shard s logically owns the INTERLEAVED rows {g : g % W == s}
(sharding.py:185-186).  The chief's merge (training/saver.py:89-180) therefore
produces a table whose merged row `shard_offset(s) + r` holds the embedding of
logical id `r * W + s`.  These helpers reproduce exactly that layout so tables
trained here restore in the reference (and vice versa) at the same W, and make
the permutation explicit for anything that wants logical row order.
Host-side utilities on CPU/GPU tensors (not part of the per-step path).
"""
import torch

from hybridbackend_b200.embedding.sharding import shard_offset
from hybridbackend_b200.embedding.sharding import shard_rows


def merge_shards(parts, bucket_size):
  """parts[s]: shard s `[shard_rows(s), dim]` -> the reference's merged variable."""
  W = len(parts)
  dim = parts[0].shape[1]
  full = torch.empty(bucket_size, dim, dtype=parts[0].dtype, device=parts[0].device)
  for s, p in enumerate(parts):
    rows = shard_rows(bucket_size, W, s)
    if p.shape[0] != rows:
      raise ValueError(f'part {s} has {p.shape[0]} rows, expected {rows}')
    off = shard_offset(bucket_size, W, s)
    full[off:off + rows] = p
  return full


def split_merged(full, num_shards):
  """Inverse of merge_shards (what restoring a merged checkpoint at the same W does)."""
  n = full.shape[0]
  return [full[shard_offset(n, num_shards, s):shard_offset(n, num_shards, s) + shard_rows(n, num_shards, s)]
          for s in range(num_shards)]


def logical_rows_of_merged(bucket_size, num_shards):
  """perm[m] = logical id stored at merged row m;  merged[inv(perm)] is the table
  in logical id order (what an unsharded PREDICT-mode graph indexes,
  training/saved_model.py:92 with embedding/sharding.py:72-75)."""
  perm = torch.empty(bucket_size, dtype=torch.int64)
  for s in range(num_shards):
    rows = shard_rows(bucket_size, num_shards, s)
    off = shard_offset(bucket_size, num_shards, s)
    perm[off:off + rows] = torch.arange(rows, dtype=torch.int64) * num_shards + s
  return perm
