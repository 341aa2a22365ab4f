"""Host side of the fused sharded GroupLookup (hbSharded* in include/hb_b200.h)."""
import ctypes as C

import torch

from hybridbackend_b200 import _lib
from hybridbackend_b200 import _util


def plan_window_bytes(world, max_nnz, dims, capacity_factor=None):
  cf = float(capacity_factor or world)
  n = len(max_nnz)
  if n == 0:
    return 1 << 20
  return int(_lib.lib().hbShardedPlanWindowBytes(
      world, n, _lib.i64_array(max_nnz), _lib.i32_array(dims), C.c_double(cf)))


def group_window_bytes(world, sizes, dims, max_nnz, capacity_factor=None, batch_size=-1):
  """Window size a Collective needs to host a GroupLookup over tables of `sizes` rows:
  the sharded plan of the tables the sharding rule shards, plus both halves of the
  dense-gradient all-reduce of the replicated small ones (hbAllreduceSumF32 needs
  world x count floats per half), plus 1 MB of slack for the op-surface collectives."""
  from hybridbackend_b200.embedding.sharding import is_small_table  # pylint: disable=import-outside-toplevel
  sh = [k for k, n in enumerate(sizes) if world > 1 and not is_small_table(n, world, batch_size)]
  plan = plan_window_bytes(world, [max_nnz[k] for k in sh], [dims[k] for k in sh], capacity_factor) if sh else 0
  dense = sum(sizes[k] * dims[k] for k in range(len(sizes)) if k not in sh) if world > 1 else 0
  ar = 2 * world * ((dense * 4 + 255) // 256 * 256 + 256)
  return int(plan + ar + (1 << 20))


class ShardedGroup:
  """The sharded features of a GroupLookup: partition -> NVSwitch push -> owner
  gather -> stitch/pool (forward) and gradient push -> owner dedup + sparse
  update (backward), all through the plan built on the Collective's window."""

  def __init__(self, collective, tables, combiners, max_nnz, capacity_factor=None):
    self.coll = collective
    self.tables = tables
    self.combiners = combiners
    self.n = len(tables)
    self.max_nnz = [int(m) for m in max_nnz]
    self.dims = [t.dim for t in tables]
    cf = float(capacity_factor or collective.world_size)
    need = plan_window_bytes(collective.world_size, self.max_nnz, self.dims, cf)
    if collective.window_bytes < need:
      raise RuntimeError(f'Collective window {collective.window_bytes} B < {need} B needed by '
                         'the sharded plan (create the Collective with window_bytes from '
                         'embedding.sharded.plan_window_bytes)')
    self._plan = C.c_void_p()
    with torch.cuda.device(collective.device):
      _lib.check(_lib.lib().hbShardedPlanCreate(
          collective.handle, self.n, _lib.i64_array(self.max_nnz), _lib.i32_array(self.dims),
          C.c_double(cf), C.byref(self._plan)), 'ShardedGroup')
    self._saved = None

  def _feats(self, ids, offsets, B, out=None, out_cols=None, grad=None, grad_cols=None,
             optimizer=None):
    feats = (_lib.hbShardedFeature * self.n)()
    for j in range(self.n):
      t = self.tables[j]
      slots = t.ensure_slots(optimizer) if optimizer is not None else t.slots
      feats[j] = _lib.hbShardedFeature(
          t.weight.data_ptr(), slots[0].data_ptr() if len(slots) > 0 else None,
          slots[1].data_ptr() if len(slots) > 1 else None, t.rows, ids[j].data_ptr(),
          offsets[j].data_ptr() if offsets[j] is not None else None, B, ids[j].numel(),
          out[:, out_cols[j]:].data_ptr() if out is not None else None,
          out.stride(0) if out is not None else 0,
          grad[:, grad_cols[j]:].data_ptr() if grad is not None else None,
          grad.stride(0) if grad is not None else 0, t.dim, _lib.COMBINER[self.combiners[j]])
    return feats

  def forward(self, ids, offsets, B, out, out_cols, status):
    for j in range(self.n):
      if ids[j].numel() > self.max_nnz[j]:
        raise ValueError(f'sharded feature {j}: nnz {ids[j].numel()} exceeds max_nnz {self.max_nnz[j]}')
    feats = self._feats(ids, offsets, B, out=out, out_cols=out_cols)
    _lib.check(_lib.lib().hbShardedLookupForward(self._plan, feats, C.c_void_p(status.data_ptr()),
                                                 _util.stream_ptr()), 'sharded forward')
    self._saved = (ids, offsets, B)

  def backward_update(self, grad, grad_cols, optimizer, desc, status):
    ids, offsets, B = self._saved
    feats = self._feats(ids, offsets, B, grad=grad, grad_cols=grad_cols, optimizer=optimizer)
    _lib.check(_lib.lib().hbShardedLookupBackwardUpdate(
        self._plan, feats, C.byref(desc), C.c_void_p(status.data_ptr()), _util.stream_ptr()),
               'sharded backward')

  def close(self):
    if self._plan:
      _lib.lib().hbShardedPlanDestroy(self._plan)
      self._plan = C.c_void_p()
