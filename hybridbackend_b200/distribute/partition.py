"""partition_by_modulo / dual modulo: host side of K1.

Mirrors hybridbackend/tensorflow/distribute/partition/ops.py:57-221 (same names,
argument meaning and return triple `(output, sizes, indices)`).  A list/tuple of
id vectors maps onto the packed HbPartitionByModuloN form
(partition_by_modulo_ops.cc:124-207): one launch sequence for all of them.
"""
import torch

from hybridbackend_b200 import _lib
from hybridbackend_b200 import _util

_ID_DTYPES = ('torch.int32', 'torch.int64', 'torch.uint32', 'torch.uint64')


def _run(ids, num_partitions, stage, modulus, opname):
  single = isinstance(ids, torch.Tensor)
  xs = [ids] if single else list(ids)
  if len(xs) < 1:
    raise ValueError(f'{opname}: N must be >= 1')
  num_partitions = int(num_partitions)
  if num_partitions < 1:
    raise ValueError(f'{opname}: num_partitions must be >= 1')
  for x in xs:
    _util.require_cuda(x, opname)
    if x.dim() != 1:
      # partition_by_modulo_ops.cc:81-83
      raise ValueError(f'{opname} expects a 1D vector.')
    if str(x.dtype) not in _ID_DTYPES:
      raise TypeError(f'{opname}: T must be one of int32,int64,uint32,uint64')
    if x.dtype != xs[0].dtype or x.device != xs[0].device:
      raise TypeError(f'{opname}: all inputs must share dtype and device')
  dev = xs[0].device
  n = len(xs)
  lens = [int(x.numel()) for x in xs]
  outs = [torch.empty_like(x) for x in xs]
  sizes = [torch.empty(num_partitions, dtype=torch.int32, device=dev) for _ in xs]
  idx = [torch.empty(l, dtype=torch.int32, device=dev) for l in lens]
  L = _lib.lib()
  need = _lib.C.c_size_t(0)
  c_lens = _lib.i32_array(lens)
  _lib.check(L.hbPartitionWorkspaceBytes(n, c_lens, num_partitions, _lib.C.byref(need)), opname)
  ws = _util.workspace(need.value, dev, 'partition')
  args = [_lib.ptr_array([x.data_ptr() for x in xs]), c_lens, _lib.C.c_int32(num_partitions)]
  tail = [_lib.ptr_array([o.data_ptr() for o in outs]),
          _lib.ptr_array([s.data_ptr() for s in sizes]),
          _lib.ptr_array([i.data_ptr() for i in idx]),
          _lib.C.c_void_p(ws.data_ptr()), _lib.C.c_size_t(ws.numel()), _util.stream_ptr()]
  with torch.cuda.device(dev):
    if stage == 0:
      rc = L.hbPartitionByModuloN(_util.dtype_code(xs[0]), n, *args, *tail)
    else:
      rc = L.hbPartitionByDualModuloN(_util.dtype_code(xs[0]), stage, n, *args,
                                      _lib.C.c_int32(int(modulus)), *tail)
  _lib.check(rc, opname)
  if single:
    return outs[0], sizes[0], idx[0]
  return outs, sizes, idx


def partition_by_modulo(ids, num_partitions, name=None):
  """Shuffle IDs using the floormod strategy (partition/ops.py:86-103).

  Returns (output, sizes, indices): ids grouped by `id mod num_partitions`
  (stable), the size of each shard, and indices for gathering back
  (`output[indices] == ids`)."""
  del name
  return _run(ids, num_partitions, 0, 1, 'partition_by_modulo')


def partition_by_dual_modulo_stage_one(ids, num_partitions, modulus, name=None):
  """Two-staged (local then global modulo) shuffle, stage one
  (partition/ops.py:106-163)."""
  del name
  return _run(ids, num_partitions, 1, modulus, 'partition_by_dual_modulo_stage_one')


def partition_by_dual_modulo_stage_two(ids, num_partitions, modulus, name=None):
  """Stage two of the dual modulo shuffle (partition/ops.py:166-221)."""
  del name
  return _run(ids, num_partitions, 2, modulus, 'partition_by_dual_modulo_stage_two')
