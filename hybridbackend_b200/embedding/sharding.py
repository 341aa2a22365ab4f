"""Table sharding rule of hybridbackend/tensorflow/embedding/variables.py:77-146.

Shard s of an N-row table over W ranks holds the global rows {g : g % W == s}
at local row g // W (embedding/sharding.py:185-186); tables with N <= W or
N <= batch_size stay replicated ("small", variables.py:95-105).
"""
import torch


def shard_rows(bucket_size, num_shards, shard):
  rows = bucket_size // num_shards           # variables.py:107
  if shard < bucket_size % num_shards:       # :108-109
    rows += 1
  return rows


def shard_offset(bucket_size, num_shards, shard):
  """SaveSliceInfo row offset of a shard (variables.py:118-123)."""
  off = (bucket_size // num_shards) * shard
  rem = bucket_size % num_shards
  return off + (shard if shard < rem else rem)


def is_small_table(bucket_size, num_shards, batch_size=-1):
  return bucket_size <= num_shards or bucket_size <= batch_size  # variables.py:96


class ShardedEmbeddingWeights:
  """One rank's part of an embedding table plus its optimizer slots.

  name follows the reference ('<name>/part_<rank>', variables.py:113-114);
  `sharded` is False for small tables (kept whole on every rank)."""

  def __init__(self, name, bucket_size, dim, rank=0, world_size=1, batch_size=-1,
               device='cuda', initializer=None, dtype=torch.float32):
    self.bucket_size = int(bucket_size)
    self.dim = int(dim)
    self.rank, self.world_size = int(rank), int(world_size)
    self.sharded = world_size > 1 and not is_small_table(bucket_size, world_size, batch_size)
    if self.sharded:
      self.rows = shard_rows(bucket_size, world_size, rank)
      self.name = f'{name}/part_{rank}'
      self.save_slice_offset = shard_offset(bucket_size, world_size, rank)
    else:
      self.rows = self.bucket_size
      self.name = name
      self.save_slice_offset = 0
    self.weight = torch.empty(self.rows, self.dim, dtype=dtype, device=device)
    if initializer is not None:
      initializer(self.weight)
    self.slots = []

  def ensure_slots(self, optimizer):
    while len(self.slots) < optimizer.num_slots:
      k = len(self.slots)
      self.slots.append(torch.full_like(self.weight, optimizer.slot_init(k)))
    return self.slots

  def load_global(self, full_table):
    """Fill this part from a full [bucket_size, dim] table (row-interleaved)."""
    if self.sharded:
      self.weight.copy_(full_table[self.rank::self.world_size].to(self.weight.device))
    else:
      self.weight.copy_(full_table.to(self.weight.device))
