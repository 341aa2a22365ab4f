"""Slab-hash cache probe: host side of HbLookup.

Mirrors the `HbLookup` op (hybridbackend/tensorflow/embedding/lookup_ops.cc:38-145):
  lookup(keys_cache, keys, cache_slab_size=32)
    -> hit_keys_indices   int32  indices into `keys` found in the cache
       hit_cache_indices  int64  their positions in `keys_cache`
       miss_keys_indices  int32  indices into `keys` not found
       miss_keys          int64  the missing keys
`keys_cache` is `slabs * 32` int64 entries, empty slot = INT64_MIN, probed
linearly from slab `murmur3_hash32(key) % slabs`.  Like the reference op, the
call reads the miss count back to the host to shape its outputs
(lookup_ops.cc:121-124).  The reference slices the miss outputs as
`[miss_count, key_count)` (lookup_ops.cc:129-131), which drops or mixes entries
unless exactly half the keys miss; this implementation returns the evidently
intended `[key_count - miss_count, key_count)`.
"""
import torch

from hybridbackend_b200 import _lib
from hybridbackend_b200 import _util

WARP_SIZE = 32


def lookup(keys_cache, keys, cache_slab_size=WARP_SIZE):
  _util.require_cuda(keys_cache, 'lookup: keys_cache')
  _util.require_cuda(keys, 'lookup: keys')
  if keys_cache.dtype != torch.int64 or keys.dtype != torch.int64:
    raise TypeError('lookup: keys_cache and keys must be int64 (the only registered kernel)')
  if keys_cache.dim() != 1:
    raise ValueError('keys_cache expects a 1D vector.')   # lookup_ops.cc:72-73
  if keys.dim() != 1:
    raise ValueError('keys expects a 1D vector.')         # lookup_ops.cc:79-80
  if cache_slab_size != WARP_SIZE:
    raise ValueError('cache_slab_size must be 32 (one slab per warp-wide probe)')
  n = keys.numel()
  slabs = keys_cache.numel() // cache_slab_size
  dev = keys.device
  if n == 0:
    e32 = torch.empty(0, dtype=torch.int32, device=dev)
    e64 = torch.empty(0, dtype=torch.int64, device=dev)
    return e32, e64, e32.clone(), e64.clone()
  if slabs < 1:
    raise ValueError('keys_cache must hold at least one slab of 32 keys')
  idx = torch.empty(n, dtype=torch.int32, device=dev)
  pay = torch.empty(n, dtype=torch.int64, device=dev)
  cnt = torch.empty(2, dtype=torch.int32, device=dev)
  with torch.cuda.device(dev):
    _lib.check(_lib.lib().hbCacheLookup(
        _lib.C.c_void_p(keys_cache.data_ptr()), _lib.C.c_int64(slabs),
        _lib.C.c_void_p(keys.data_ptr()), _lib.C.c_int32(n), _lib.C.c_void_p(idx.data_ptr()),
        _lib.C.c_void_p(pay.data_ptr()), _lib.C.c_void_p(cnt.data_ptr()), _util.stream_ptr()),
               'lookup')
  miss = int(cnt[0].item())  # host read, as the reference (BlockHostUntilDone)
  return idx[:n - miss], pay[:n - miss], idx[n - miss:], pay[n - miss:]
