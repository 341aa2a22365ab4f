"""ctypes/numpy front-end of the CPU oracle (oracle/hb_oracle.c) and, when built,
of the reference's own CPU functors (oracle/_ref/libhbref.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
hybridbackend_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DT = {np.dtype(np.int32): 0, np.dtype(np.int64): 1,
       np.dtype(np.uint32): 2, np.dtype(np.uint64): 3}
COMBINER = {'sum': 0, 'mean': 1, 'sqrtn': 2}


def build(force=False):
  so = os.path.join(_HERE, 'libhb_oracle.so')
  src = os.path.join(_HERE, 'hb_oracle.c')
  srcs = [src, os.path.join(_HERE, 'hb_oracle_mt.c'), os.path.join(_HERE, 'hb_oracle.h')]
  if force or not os.path.exists(so) or any(
      os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
    subprocess.check_call(['make', '-C', _HERE, 'libhb_oracle.so'],
                          stdout=subprocess.DEVNULL)
  if os.path.isdir('/root/reference/hybridbackend'):
    subprocess.check_call([os.path.join(_HERE, 'build_ref.sh')],
                          stdout=subprocess.DEVNULL)
  return so


_lib = None
_ref = None


def lib():
  global _lib
  if _lib is None:
    _lib = C.CDLL(build())
    _lib.hbo_shard_rows.restype = C.c_int64
    _lib.hbo_shard_offset.restype = C.c_int64
    _lib.hbo_unique_i64.restype = C.c_int64
    _lib.hbo_cache_lookup.restype = C.c_int64
    _lib.hbo_murmur3_hash32_i64.restype = C.c_uint32
    _lib.hbo_shard_rows.argtypes = [C.c_int64, C.c_int, C.c_int]
    _lib.hbo_shard_offset.argtypes = [C.c_int64, C.c_int, C.c_int]
    _lib.hbo_is_small_table.argtypes = [C.c_int64, C.c_int, C.c_int64]
    _lib.hbo_murmur3_hash32_i64.argtypes = [C.c_int64]
  return _lib


def ref():
  """The reference's own functors (None when oracle/_ref was never built)."""
  global _ref
  if _ref is None:
    so = os.path.join(_HERE, '_ref', 'libhbref.so')
    if not os.path.exists(so):
      build()
    if not os.path.exists(so):
      return None
    _ref = C.CDLL(so)
    _ref.hbref_murmur3_hash32_i64.restype = C.c_uint32
    _ref.hbref_murmur3_hash32_i64.argtypes = [C.c_longlong]
  return _ref


def _p(a):
  return a.ctypes.data_as(C.c_void_p)


def _check(rc, what):
  if rc != 0:
    raise RuntimeError(f'oracle {what} failed rc={rc}')


def _partition(handle, fn, x, p, extra=()):
  x = np.ascontiguousarray(x)
  n = x.shape[0]
  out = np.empty_like(x)
  sizes = np.zeros(p, np.int32)
  idx = np.empty(n, np.int32)
  rc = getattr(handle, fn)(_DT[x.dtype], *extra[:1], _p(x), C.c_int32(n),
                           C.c_int32(p), *extra[1:], _p(out), _p(sizes), _p(idx))
  _check(rc, fn)
  return out, sizes, idx


def partition_by_modulo(x, num_partitions):
  return _partition(lib(), 'hbo_partition_by_modulo', x, num_partitions)


def partition_by_dual_modulo(x, num_partitions, modulus, stage):
  return _partition(lib(), 'hbo_partition_by_dual_modulo', x, num_partitions,
                    (C.c_int(stage), C.c_int32(modulus)))


def ref_partition_by_modulo(x, num_partitions):
  return _partition(ref(), 'hbref_partition_by_modulo', x, num_partitions)


def ref_partition_by_dual_modulo(x, num_partitions, modulus, stage):
  return _partition(ref(), 'hbref_partition_by_dual_modulo', x, num_partitions,
                    (C.c_int(stage), C.c_int32(modulus)))


def alltoallv(inputs, sizes, common_shape=()):
  """inputs[r]: array [sum(sizes[r]), *common_shape]; sizes[r]: W ints.
  Returns (outputs, out_sizes) per rank (nccl_collective.cc:250-288)."""
  W = len(inputs)
  inputs = [np.ascontiguousarray(a) for a in inputs]
  dt = inputs[0].dtype
  common = int(np.prod(common_shape)) if len(common_shape) else 1
  s = np.ascontiguousarray(np.asarray(sizes, np.int32).reshape(W, W))
  tot = s.sum(axis=0)
  outs = [np.empty((int(tot[r]),) + tuple(common_shape), dt) for r in range(W)]
  rs = np.zeros((W, W), np.int32)
  send = (C.c_void_p * W)(*[a.ctypes.data for a in inputs])
  recv = (C.c_void_p * W)(*[a.ctypes.data for a in outs])
  rc = lib().hbo_alltoallv(W, send, _p(s), C.c_int64(common),
                           C.c_int(dt.itemsize), recv, _p(rs))
  _check(rc, 'alltoallv')
  return outs, [rs[r].copy() for r in range(W)]


def alltoallv_n(n_inputs, n_sizes, n_common_shape=None):
  """n_inputs[r][k], n_sizes[r][k] -> per rank list over k (AlltoallvN)."""
  W = len(n_inputs)
  N = len(n_inputs[0])
  res = [[None] * N for _ in range(W)]
  for k in range(N):
    cs = () if n_common_shape is None else tuple(n_common_shape[k])
    outs, rs = alltoallv([n_inputs[r][k] for r in range(W)],
                         [n_sizes[r][k] for r in range(W)], cs)
    for r in range(W):
      res[r][k] = (outs[r], rs[r])
  return res


def shard_rows(n, w, s):
  return lib().hbo_shard_rows(n, w, s)


def shard_offset(n, w, s):
  return lib().hbo_shard_offset(n, w, s)


def is_small_table(n, w, batch_size=-1):
  return bool(lib().hbo_is_small_table(n, w, batch_size))


def unique(ids):
  ids = np.ascontiguousarray(ids, np.int64)
  uniq = np.empty_like(ids)
  inv = np.empty(ids.shape[0], np.int32)
  u = lib().hbo_unique_i64(_p(ids), C.c_int64(ids.shape[0]), _p(uniq), _p(inv))
  return uniq[:u].copy(), inv


def embedding_lookup_sparse(table, ids, offsets, combiner='mean', out=None,
                            dedup=True):
  table = np.ascontiguousarray(table, np.float32)
  ids = np.ascontiguousarray(ids, np.int64)
  offsets = np.ascontiguousarray(offsets, np.int64)
  nb = offsets.shape[0] - 1
  dim = table.shape[1]
  if out is None:
    out = np.empty((nb, dim), np.float32)
  fn = lib().hbo_embedding_lookup_sparse if dedup else lib().hbo_embedding_bag
  rc = fn(_p(table), C.c_int64(table.shape[0]), C.c_int(dim), _p(ids),
          _p(offsets), C.c_int64(nb), C.c_int(COMBINER[combiner]), _p(out),
          C.c_int64(out.strides[0] // 4))
  _check(rc, 'embedding_lookup_sparse')
  return out


def lookup_row_grads(grad, offsets, combiner='mean'):
  grad = np.asarray(grad, np.float32)
  assert grad.strides[1] == 4
  offsets = np.ascontiguousarray(offsets, np.int64)
  nb = offsets.shape[0] - 1
  dim = grad.shape[1]
  rg = np.empty((int(offsets[-1]), dim), np.float32)
  rc = lib().hbo_lookup_row_grads(_p(grad), C.c_int64(grad.strides[0] // 4),
                                  C.c_int(dim), _p(offsets), C.c_int64(nb),
                                  C.c_int(COMBINER[combiner]), _p(rg))
  _check(rc, 'lookup_row_grads')
  return rg


def sparse_apply_adagrad(table, accum, rows_idx, row_grad, lr):
  """In place on table/accum (float32, C-contiguous)."""
  rows_idx = np.ascontiguousarray(rows_idx, np.int64)
  row_grad = np.ascontiguousarray(row_grad, np.float32)
  rc = lib().hbo_sparse_apply_adagrad(
      _p(table), _p(accum), C.c_int64(table.shape[0]), C.c_int(table.shape[1]),
      _p(rows_idx), _p(row_grad), C.c_int64(rows_idx.shape[0]), C.c_float(lr))
  _check(rc, 'sparse_apply_adagrad')


def sparse_apply_lazy_adam(table, m, v, rows_idx, row_grad, lr, beta1=0.9,
                           beta2=0.999, eps=1e-8, step=1):
  rows_idx = np.ascontiguousarray(rows_idx, np.int64)
  row_grad = np.ascontiguousarray(row_grad, np.float32)
  rc = lib().hbo_sparse_apply_lazy_adam(
      _p(table), _p(m), _p(v), C.c_int64(table.shape[0]),
      C.c_int(table.shape[1]), _p(rows_idx), _p(row_grad),
      C.c_int64(rows_idx.shape[0]), C.c_float(lr), C.c_float(beta1),
      C.c_float(beta2), C.c_float(eps), C.c_int64(step))
  _check(rc, 'sparse_apply_lazy_adam')


def shard_table(table, w):
  """Row-interleaved shards: shard s holds global rows g with g % w == s at
  local row g // w (embedding/sharding.py:185-186, variables.py:107-111)."""
  return [np.ascontiguousarray(table[s::w]) for s in range(w)]


def sharded_embedding_lookup(shards, bucket_size, ids_per_rank):
  W = len(shards)
  shards = [np.ascontiguousarray(s, np.float32) for s in shards]
  dim = shards[0].shape[1]
  ids = [np.ascontiguousarray(i, np.int64) for i in ids_per_rank]
  n = np.asarray([i.shape[0] for i in ids], np.int64)
  outs = [np.empty((int(n[r]), dim), np.float32) for r in range(W)]
  sp = (C.c_void_p * W)(*[a.ctypes.data for a in shards])
  ip = (C.c_void_p * W)(*[a.ctypes.data for a in ids])
  op = (C.c_void_p * W)(*[a.ctypes.data for a in outs])
  rc = lib().hbo_sharded_embedding_lookup(W, sp, C.c_int64(bucket_size),
                                          C.c_int(dim), ip, _p(n), op)
  _check(rc, 'sharded_embedding_lookup')
  return outs


def murmur3_hash32(key):
  return int(lib().hbo_murmur3_hash32_i64(C.c_int64(int(key))))


def ref_murmur3_hash32(key):
  return int(ref().hbref_murmur3_hash32_i64(C.c_longlong(int(key))))


def cache_lookup(keys_cache, keys):
  keys_cache = np.ascontiguousarray(keys_cache, np.int64)
  keys = np.ascontiguousarray(keys, np.int64)
  n = keys.shape[0]
  slabs = keys_cache.shape[0] // 32
  hi = np.empty(n, np.int32); hc = np.empty(n, np.int64)
  mi = np.empty(n, np.int32); mk = np.empty(n, np.int64)
  nh = C.c_int64(0)
  nm = lib().hbo_cache_lookup(_p(keys_cache), C.c_int64(slabs), _p(keys),
                              C.c_int64(n), _p(hi), _p(hc), _p(mi), _p(mk),
                              C.byref(nh))
  return hi[:nh.value], hc[:nh.value], mi[:nm], mk[:nm]


class _MtFeature(C.Structure):
  _fields_ = [('table', C.c_void_p), ('accum', C.c_void_p), ('rows', C.c_int64), ('dim', C.c_int32),
              ('parts', C.c_int32), ('ids', C.c_void_p), ('nbags', C.c_int64), ('grad', C.c_void_p),
              ('grad_stride', C.c_int64), ('out', C.c_void_p), ('out_stride', C.c_int64)]


def mt_step(tables, accums, ids, grad, out, lr, nthreads, fwd_chunk=8192, parts=None):
  """One forward(+backward) step over F one-id-per-bag features on `nthreads`
  pthreads (hbo_mt_step).  tables/accums: lists of [rows, dim] float32 arrays
  (accums None: forward only); ids: list of int64 [B]; grad/out: [B, F*dim]."""
  F = len(tables)
  dim = tables[0].shape[1]
  feats = (_MtFeature * F)()
  for k in range(F):
    g = grad[:, k * dim:(k + 1) * dim]
    oo = out[:, k * dim:(k + 1) * dim]
    feats[k] = _MtFeature(tables[k].ctypes.data, accums[k].ctypes.data if accums is not None else None,
                          tables[k].shape[0], dim, 1 if parts is None else int(parts[k]),
                          ids[k].ctypes.data, ids[k].shape[0], g.ctypes.data, grad.strides[0] // 4,
                          oo.ctypes.data, out.strides[0] // 4)
  rc = lib().hbo_mt_step(F, feats, C.c_int64(fwd_chunk), C.c_float(lr), int(nthreads))
  _check(rc, 'mt_step')
