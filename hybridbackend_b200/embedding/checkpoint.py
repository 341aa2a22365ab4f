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
import json
import os

import numpy as np
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


# ---------------------------------------------------------------------------------------------
# Save / merge / restore protocol of sharded tables (training/saver.py:89-180), in a neutral
# container.  The reference's saver has every rank write its `<name>/part_<rank>` slices
# (with their SaveSliceInfo) to a temporary prefix, waits on a barrier, lets the chief merge
# the per-rank files into ONE checkpoint under the user's prefix and deletes the temporaries;
# restoring at the same world size hands every rank its slice of the merged variable.  TF's
# tensor-bundle file format cannot be produced without TensorFlow, so the files here are
# `.npz` archives carrying the same variable names, slice specs and the same merged layout
# (merge_shards above); a TF-side converter only has to copy arrays.
# ---------------------------------------------------------------------------------------------
def slice_info(name, bucket_size, dim, num_shards, rank):
  """The SaveSliceInfo of `<name>/part_<rank>` (embedding/variables.py:118-132):
  full_name, full_shape, var_offset, var_shape."""
  return {'full_name': name, 'full_shape': [int(bucket_size), int(dim)],
          'var_offset': [int(shard_offset(bucket_size, num_shards, rank)), 0],
          'var_shape': [int(shard_rows(bucket_size, num_shards, rank)), int(dim)]}


def _part_file(prefix, tag, rank):
  return f'{prefix}_temp_{tag}/part-{rank:05d}.npz'


def save_local_shards(prefix, tag, rank, num_shards, tables, slots=None):
  """Rank `rank` writes its slices: tables = {name: (bucket_size, shard [rows, dim])};
  slots = {name: {slot_name: shard}} (optimizer slots are sharded like their variable,
  training/optimizer.py:102-118).  `tag` is the checkpoint uuid all ranks agreed on.
  Returns the file name (the rank then signals the saver's local barrier)."""
  arrays, specs = {}, {}

  def put(full, bucket_size, part):
    part = np.ascontiguousarray(part.detach().cpu().numpy() if hasattr(part, 'detach') else part)
    info = slice_info(full, bucket_size, part.shape[1], num_shards, rank)
    if list(part.shape) != info['var_shape']:
      raise ValueError(f'{full}/part_{rank}: shape {list(part.shape)} != {info["var_shape"]}')
    key = f'{full}/part_{rank}'
    arrays[key] = part
    specs[key] = info
  for name, (bucket_size, part) in tables.items():
    put(name, bucket_size, part)
    for slot_name, sp in (slots or {}).get(name, {}).items():
      put(f'{name}/{slot_name}', bucket_size, sp)
  path = _part_file(prefix, tag, rank)
  os.makedirs(os.path.dirname(path), exist_ok=True)
  tmp = path + '.tmp'
  with open(tmp, 'wb') as f:
    np.savez(f, __specs__=np.frombuffer(json.dumps(specs).encode(), np.uint8), **arrays)
  os.replace(tmp, path)
  return path


def merge_checkpoint(prefix, tag, num_shards, delete_old_dirs=True):
  """The chief's MergeV2Checkpoints step: all `part-*.npz` of the temporary prefix become
  `<prefix>.npz` holding every variable in the reference's merged layout."""
  merged, specs_all = {}, {}
  for rank in range(num_shards):
    path = _part_file(prefix, tag, rank)
    if not os.path.exists(path):
      raise FileNotFoundError(f'rank {rank} has not written {path} (barrier not passed?)')
    with np.load(path) as z:
      specs = json.loads(bytes(z['__specs__']).decode())
      for key, info in specs.items():
        full = merged.get(info['full_name'])
        if full is None:
          full = np.empty(info['full_shape'], z[key].dtype)
          merged[info['full_name']] = full
          specs_all[info['full_name']] = {'full_shape': info['full_shape'], 'num_shards': num_shards}
        off, rows = info['var_offset'][0], info['var_shape'][0]
        full[off:off + rows] = z[key]
  out = prefix + '.npz'
  tmp = out + '.tmp'
  with open(tmp, 'wb') as f:
    np.savez(f, __specs__=np.frombuffer(json.dumps(specs_all).encode(), np.uint8), **merged)
  os.replace(tmp, out)
  if delete_old_dirs:
    for rank in range(num_shards):
      os.remove(_part_file(prefix, tag, rank))
    os.rmdir(os.path.dirname(_part_file(prefix, tag, 0)))
  return out


def restore_local_shards(prefix, rank, num_shards, names=None):
  """What restoring the merged checkpoint at the same world size gives rank `rank`:
  {full_name: its slice [shard_rows, dim]} (variables and slots alike)."""
  out = {}
  with np.load(prefix + '.npz') as z:
    specs = json.loads(bytes(z['__specs__']).decode())
    for full, info in specs.items():
      if names is not None and full not in names:
        continue
      if info['num_shards'] != num_shards:
        raise ValueError(f'{full} was saved with {info["num_shards"]} shards; the merged row order is '
                         f'only meaningful at the same world size (see logical_rows_of_merged)')
      n = info['full_shape'][0]
      off, rows = shard_offset(n, num_shards, rank), shard_rows(n, num_shards, rank)
      out[full] = torch.from_numpy(np.array(z[full][off:off + rows]))
  return out
