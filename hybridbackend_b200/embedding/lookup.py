"""embedding_lookup_sparse / GroupLookup: host side of K3 (gather+pool), K5
(backward + sparse optimizer) and of the fused sharded path (K1+K2+K3+K4).

Mirrors what `tf.nn.embedding_lookup_sparse` computes under
`hb.embedding_scope()` (hybridbackend/tensorflow/embedding/sharding.py:168-205):
sparse ids are given in CSR form (values + bag offsets; a TF SparseTensor's
`indices[:, 0]` converts with `segment_ids_to_offsets`), `sp_weights` is None and
the default combiner is "mean" as in TF-1.15.
"""
import torch

from hybridbackend_b200 import _lib
from hybridbackend_b200 import _util
from hybridbackend_b200.embedding.sharding import ShardedEmbeddingWeights


def segment_ids_to_offsets(segment_ids, nbags):
  """Sorted int segment ids [nnz] (SparseTensor.indices[:, 0]) -> offsets
  [nbags+1] int64.  Data-prep convenience (torch ops), not part of the hot path."""
  seg = segment_ids.to(torch.int64)
  bounds = torch.arange(nbags + 1, device=seg.device, dtype=torch.int64)
  return torch.searchsorted(seg, bounds, right=False).to(torch.int64)


def _weight_of(params):
  return params.weight if isinstance(params, ShardedEmbeddingWeights) else params


def _feature_struct(weight, ids, offsets, nbags, out, out_stride, combiner, id_div=1):
  return _lib.hbLookupFeature(
      weight.data_ptr(), weight.shape[0], ids.data_ptr(),
      offsets.data_ptr() if offsets is not None else None, nbags, out.data_ptr(),
      out_stride, weight.shape[1], _lib.COMBINER[combiner], id_div, ids.numel())


def _check_inputs(weight, ids, offsets, what):
  _util.require_cuda(weight, f'{what}: params')
  _util.require_cuda(ids, f'{what}: ids')
  if weight.dtype != torch.float32 or weight.dim() != 2:
    raise TypeError(f'{what}: params must be a 2-D float32 tensor')
  if ids.dtype != torch.int64 or ids.dim() != 1:
    raise TypeError(f'{what}: ids must be a 1-D int64 tensor')
  if offsets is not None:
    _util.require_cuda(offsets, f'{what}: offsets')
    if offsets.dtype != torch.int64 or offsets.dim() != 1 or offsets.numel() < 1:
      raise TypeError(f'{what}: offsets must be a 1-D int64 tensor of nbags+1 entries')


def embedding_lookup_sparse(params, ids, offsets=None, sp_weights=None, combiner=None,
                            out=None, check=False):
  """Pooled lookup of one feature: out[b] = combine(params[ids[offsets[b]:offsets[b+1]]]).

  offsets=None means one id per bag.  combiner defaults to "mean" (TF-1.15)."""
  if sp_weights is not None:
    raise NotImplementedError('sp_weights is not on the hot path (the reference '
                              'call sites pass None)')
  combiner = combiner or 'mean'
  if combiner not in _lib.COMBINER:
    raise ValueError('combiner must be one of "mean", "sqrtn" or "sum"')
  weight = _weight_of(params)
  _check_inputs(weight, ids, offsets, 'embedding_lookup_sparse')
  nbags = ids.numel() if offsets is None else offsets.numel() - 1
  dim = weight.shape[1]
  if out is None:
    out = torch.empty(nbags, dim, dtype=torch.float32, device=weight.device)
  feat = (_lib.hbLookupFeature * 1)(
      _feature_struct(weight, ids, offsets, nbags, out, out.stride(0), combiner))
  st = _util.status_word(weight.device)
  with torch.cuda.device(weight.device):
    _lib.check(_lib.lib().hbGroupLookupForward(1, feat, _lib.C.c_void_p(st.data_ptr()),
                                               _util.stream_ptr()),
               'embedding_lookup_sparse')
  if check:
    _util.check_status(weight.device)
  return out


def embedding_lookup(params, ids, check=False):
  """tf.nn.embedding_lookup on a local table: out[i] = params[ids[i]]."""
  return embedding_lookup_sparse(params, ids.reshape(-1), None, combiner='sum', check=check)


class GroupLookup:
  """All sparse features of a model looked up (and updated) together.

  Single rank / replicated tables: one fused gather+pool launch writes straight
  into the concatenated `[B, sum(dim)]` dense-MLP input; `backward_update` sorts
  the ids once and applies the sparse optimizer in one fused pass per row.
  With a `Collective` (world_size > 1) sharded tables go through the fused
  partition -> NVSwitch push -> owner gather -> stitch path
  (embedding/sharding.py:171-203 composition; see hb_b200.h).
  """

  def __init__(self, tables, combiners=None, collective=None, max_nnz=None,
               capacity_factor=None, sync_replicated=True, overlap_backward_sort=True):
    self.tables = list(tables)
    self.n = len(self.tables)
    if self.n < 1:
      raise ValueError('GroupLookup needs at least one table')
    self.combiners = list(combiners) if combiners is not None else ['mean'] * self.n
    self.dims = [_weight_of(t).shape[1] for t in self.tables]
    self.col_offsets = [0]
    for d in self.dims:
      self.col_offsets.append(self.col_offsets[-1] + d)
    self.out_dim = self.col_offsets[-1]
    self.device = _weight_of(self.tables[0]).device
    self.collective = collective
    self.sync_replicated = sync_replicated
    # training mode: the id sort of the backward (needs only ids) is enqueued on a
    # side stream at forward time and overlaps the forward gather
    self.overlap_backward_sort = overlap_backward_sort
    self._side = None
    self._sort_done = None
    self._upd_ws = None
    self._saved = None
    self._sharded = None
    self._dense = {}  # replicated tables: dense gradient buffers and row ids
    # descriptor cache: building ~80 ctypes structs per step costs more host time than
    # the kernels take on the device, so struct arrays are kept per set of buffers
    self._cache = {}
    self._d2h_events = {}
    self._ev_ready = None
    world = collective.world_size if collective is not None else 1
    self.sharded_idx = [k for k, t in enumerate(self.tables)
                        if world > 1 and isinstance(t, ShardedEmbeddingWeights) and t.sharded]
    self.local_idx = [k for k in range(self.n) if k not in self.sharded_idx]
    if self.sharded_idx:
      from hybridbackend_b200.embedding.sharded import ShardedGroup  # pylint: disable=import-outside-toplevel
      if max_nnz is None:
        raise ValueError('GroupLookup with sharded tables needs max_nnz (static bound '
                         'of ids per feature per rank)')
      self._sharded = ShardedGroup(
          collective, [self.tables[k] for k in self.sharded_idx],
          [self.combiners[k] for k in self.sharded_idx],
          [max_nnz[k] for k in self.sharded_idx], capacity_factor)

  # -- forward ---------------------------------------------------------------
  def _memo(self, kind, key, build):
    c = self._cache.setdefault(kind, {})
    v = c.get(key)
    if v is None:
      if len(c) >= 16:
        c.clear()
      v = build()
      c[key] = v
    return v

  def forward(self, ids, offsets=None, out=None, check=False, prepare_backward=True):
    """ids[k]: int64 [nnz_k]; offsets[k]: int64 [B+1] or None.  Returns
    out [B, sum(dim)] float32 (feature k occupies columns col_offsets[k]:+dim)."""
    offsets = list(offsets) if offsets is not None else [None] * self.n
    B = ids[0].numel() if offsets[0] is None else offsets[0].numel() - 1
    if out is None:
      out = torch.empty(B, self.out_dim, dtype=torch.float32, device=self.device)
    in_key = (tuple(t.data_ptr() for t in ids), tuple(t.numel() for t in ids),
              tuple(t.data_ptr() if t is not None else 0 for t in offsets), B)

    def validate():
      nbags = [ids[k].numel() if offsets[k] is None else offsets[k].numel() - 1 for k in range(self.n)]
      if any(b != B for b in nbags):
        raise ValueError('all features of a GroupLookup must have the same number of bags')
      for k in range(self.n):
        _check_inputs(_weight_of(self.tables[k]), ids[k], offsets[k], f'GroupLookup feature {k}')
      return True
    self._memo('valid', in_key, validate)
    st = _util.status_word(self.device)
    L = _lib.lib()
    with torch.cuda.device(self.device):
      self._sort_done = None
      if self.local_idx and prepare_backward and self.overlap_backward_sort:
        self._presort(ids, offsets, B, st, in_key)
      if self.local_idx:
        def build():
          feats = (_lib.hbLookupFeature * len(self.local_idx))()
          for j, k in enumerate(self.local_idx):
            feats[j] = _feature_struct(_weight_of(self.tables[k]), ids[k], offsets[k], B,
                                       out[:, self.col_offsets[k]:], out.stride(0),
                                       self.combiners[k])
          return feats
        feats = self._memo('fwd', (in_key, out.data_ptr(), out.stride(0)), build)
        _lib.check(L.hbGroupLookupForward(len(self.local_idx), feats,
                                          _lib.C.c_void_p(st.data_ptr()), _util.stream_ptr()),
                   'GroupLookup.forward')
      if self._sharded is not None:
        self._sharded.forward([ids[k] for k in self.sharded_idx],
                              [offsets[k] for k in self.sharded_idx], B, out,
                              [self.col_offsets[k] for k in self.sharded_idx], st)
    self._saved = (list(ids), offsets, B, in_key)
    if check:
      _util.check_status(self.device)
    return out

  def _sync(self):
    return (self.collective is not None and self.collective.world_size > 1 and
            self.sync_replicated and bool(self.local_idx))

  def close(self):
    """Release the sharded plan (the Collective can then host another GroupLookup)."""
    if self._sharded is not None:
      self._sharded.close()
      self._sharded = None

  def _update_feats(self, ids, offsets, B, grad=None, optimizer=None, in_key=None):
    def build():
      m = len(self.local_idx)
      feats = (_lib.hbUpdateFeature * m)()
      for j, k in enumerate(self.local_idx):
        w = _weight_of(self.tables[k])
        slots = self._slots(k, optimizer) if optimizer is not None else []
        feats[j] = _lib.hbUpdateFeature(
            w.data_ptr(), slots[0].data_ptr() if len(slots) > 0 else None,
            slots[1].data_ptr() if len(slots) > 1 else None, w.shape[0], ids[k].data_ptr(),
            offsets[k].data_ptr() if offsets[k] is not None else None, B, ids[k].numel(),
            grad[:, self.col_offsets[k]:].data_ptr() if grad is not None else None,
            grad.stride(0) if grad is not None else w.shape[1], w.shape[1],
            _lib.COMBINER[self.combiners[k]], 1)
      return feats, self._workspace_bytes(feats)
    if in_key is None:
      return build()
    key = (in_key, grad.data_ptr() if grad is not None else 0, grad.stride(0) if grad is not None else 0,
           optimizer.kind if optimizer is not None else '')
    return self._memo('upd', key, build)

  def _workspace_bytes(self, feats):
    need = _lib.C.c_size_t(0)
    _lib.check(_lib.lib().hbGroupSparseUpdateWorkspaceBytes(len(feats), feats, _lib.C.byref(need)),
               'GroupLookup workspace')
    return need.value

  def _workspace(self, feats, need=None):
    if need is None:
      need = self._workspace_bytes(feats)
    if self._upd_ws is None or self._upd_ws.numel() < need:
      self._upd_ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
    return self._upd_ws

  def _presort(self, ids, offsets, B, st, in_key=None):
    """Enqueue the backward's id sort on the side stream (hbGroupSparseSort).  The ids
    stay referenced by self._saved until the next forward, and the backward waits for
    the side stream, so no allocator hand-over (record_stream) is needed."""
    if self._side is None:
      self._side = torch.cuda.Stream(device=self.device)
      self._ev_ready = torch.cuda.Event()
      self._ev_sorted = torch.cuda.Event()
    main = torch.cuda.current_stream()
    feats, need = self._update_feats(ids, offsets, B, in_key=in_key)
    ws = self._workspace(feats, need)
    self._ev_ready.record(main)
    self._side.wait_event(self._ev_ready)
    _lib.check(_lib.lib().hbGroupSparseSort(
        len(feats), feats, _lib.C.c_void_p(ws.data_ptr()), _lib.C.c_size_t(ws.numel()),
        _lib.C.c_void_p(st.data_ptr()), _lib.C.c_void_p(self._side.cuda_stream)), 'GroupLookup presort')
    self._ev_sorted.record(self._side)
    self._sort_done = self._ev_sorted

  def forward_host(self, h_ids, d_stage, out, h_out, check=False, d2h_stream=None):
    """Host-buffer forward (one id per bag): h_ids pinned int64 [n, B] is copied
    H2D into d_stage [n, B], the fused lookup runs, and `out` [B, sum(dim)]
    (contiguous, device) is copied D2H into pinned h_out -- all enqueued on the
    current stream by ONE C-ABI call (hbGroupLookupForwardHost).

    d2h_stream: copy the output back on that stream instead (behind an event), so the
    backward of this step and the H2D + forward of the next overlap the D2H; the caller
    rotates (out, h_out) pairs, and the next forward into the same `out` waits for its
    copy.  h_out is valid once d2h_stream (or the device) is synchronised."""
    if not (h_ids.is_pinned() and h_out.is_pinned()):
      raise ValueError('forward_host needs pinned host tensors')
    if h_ids.dtype != torch.int64 or h_ids.dim() != 2 or h_ids.shape[0] != self.n:
      raise TypeError('h_ids must be pinned int64 [n_features, B]')
    B = h_ids.shape[1]
    _util.require_cuda(d_stage, 'forward_host: d_stage')
    _util.require_cuda(out, 'forward_host: out')
    if d_stage.shape != h_ids.shape or d_stage.dtype != torch.int64:
      raise ValueError('d_stage must be an int64 device tensor shaped like h_ids')
    if tuple(out.shape) != (B, self.out_dim) or tuple(h_out.shape) != (B, self.out_dim):
      raise ValueError('out / h_out must be [B, sum(dim)]')
    st = _util.status_word(self.device)
    L = _lib.lib()
    ids = [d_stage[k] for k in range(self.n)]
    with torch.cuda.device(self.device):
      self._sort_done = None
      main = torch.cuda.current_stream()
      pending = self._d2h_events.pop(out.data_ptr(), None)
      if pending is not None:
        main.wait_event(pending)   # the previous copy out of this buffer
      if self._sharded is None:
        def build():
          feats = (_lib.hbLookupFeature * self.n)()
          for k in range(self.n):
            feats[k] = _feature_struct(_weight_of(self.tables[k]), ids[k], None, B,
                                       out[:, self.col_offsets[k]:], out.stride(0),
                                       self.combiners[k])
          return feats
        feats = self._memo('fwd_host', (d_stage.data_ptr(), out.data_ptr(), B), build)
        _lib.check(L.hbGroupLookupForwardHost(
            self.n, feats, _lib.C.c_void_p(h_ids.data_ptr()), _lib.C.c_void_p(d_stage.data_ptr()),
            _lib.C.c_size_t(h_ids.numel() * 8), _lib.C.c_void_p(out.data_ptr()),
            _lib.C.c_void_p(h_out.data_ptr()), _lib.C.c_size_t(out.numel() * 4 if d2h_stream is None else 0),
            _lib.C.c_void_p(st.data_ptr()), _util.stream_ptr()), 'GroupLookup.forward_host')
        self._saved = (ids, [None] * self.n, B, ('host', d_stage.data_ptr(), B))
      else:
        d_stage.copy_(h_ids, non_blocking=True)
        self.forward(ids, out=out)
        if d2h_stream is None:
          h_out.copy_(out, non_blocking=True)
      if d2h_stream is not None:
        ready = torch.cuda.Event()
        ready.record(main)
        d2h_stream.wait_event(ready)
        with torch.cuda.stream(d2h_stream):
          h_out.copy_(out, non_blocking=True)
          done = torch.cuda.Event()
          done.record(d2h_stream)
        self._d2h_events[out.data_ptr()] = done
    if check:
      _util.check_status(self.device)
    return h_out

  # -- backward + sparse optimizer apply ------------------------------------------
  def backward_update(self, grad, optimizer, check=False):
    """grad: [B, sum(dim)] upstream gradient of forward()'s output.  Applies
    `optimizer` to every table in place (slots are created on first use)."""
    if self._saved is None:
      raise RuntimeError('GroupLookup.backward_update called before forward')
    ids, offsets, B, in_key = self._saved
    _util.require_cuda(grad, 'GroupLookup.backward_update: grad')
    if grad.dtype != torch.float32 or grad.dim() != 2 or grad.shape[0] != B or \
        grad.shape[1] != self.out_dim or grad.stride(1) != 1:
      raise ValueError('grad must be float32 [B, sum(dim)] with unit inner stride')
    optimizer.step += 1
    desc = optimizer.descriptor()
    st = _util.status_word(self.device)
    L = _lib.lib()
    with torch.cuda.device(self.device):
      rep_done = None
      if self.local_idx and self._sync():
        # replicated small tables: densify -> all-reduce -> dense apply is a chain of a dozen
        # tiny launches; it runs on the side stream next to the sharded backward (different
        # tables, different window region) and is joined at the end
        if self._side is None:
          self._side = torch.cuda.Stream(device=self.device)
          self._ev_ready = torch.cuda.Event()
          self._ev_sorted = torch.cuda.Event()
        main = torch.cuda.current_stream()
        if self._sharded is not None:
          fork = torch.cuda.Event()
          fork.record(main)
          self._side.wait_event(fork)
          with torch.cuda.stream(self._side):
            self._replicated_dense_update(ids, offsets, B, grad, optimizer, desc, st)
            rep_done = torch.cuda.Event()
            rep_done.record(self._side)
        else:
          self._replicated_dense_update(ids, offsets, B, grad, optimizer, desc, st)
      elif self.local_idx:
        feats, need = self._update_feats(ids, offsets, B, grad, optimizer, in_key=in_key)
        m = len(feats)
        ws = self._workspace(feats, need)
        if self._sort_done is not None:
          torch.cuda.current_stream().wait_event(self._sort_done)
          self._sort_done = None
          _lib.check(L.hbGroupSparseApply(
              m, feats, _lib.C.byref(desc), _lib.C.c_void_p(ws.data_ptr()),
              _lib.C.c_size_t(ws.numel()), _lib.C.c_void_p(st.data_ptr()), _util.stream_ptr()),
                     'GroupLookup.backward_update')
        else:
          _lib.check(L.hbGroupLookupBackwardUpdate(
              m, feats, _lib.C.byref(desc), _lib.C.c_void_p(ws.data_ptr()),
              _lib.C.c_size_t(ws.numel()), _lib.C.c_void_p(st.data_ptr()), _util.stream_ptr()),
                     'GroupLookup.backward_update')
      if self._sharded is not None:
        self._sharded.backward_update(grad, [self.col_offsets[k] for k in self.sharded_idx],
                                      optimizer, desc, st)
      if rep_done is not None:
        torch.cuda.current_stream().wait_event(rep_done)
    if check:
      _util.check_status(self.device)

  def _replicated_dense_update(self, ids, offsets, B, grad, optimizer, desc, st):
    """Replicated ("small") tables at world_size > 1, as the reference routes them
    (training/gradient.py:132-141): the sparse gradient is DENSIFIED to [rows, dim]
    (duplicates summed in position order), all-reduced over the ranks (:157-160),
    multiplied by 1/W (_mean, :77-97 via :216) and applied as a dense gradient, i.e.
    to every row.  Device work: (1) the fused sort + duplicate-sum kernels with
    SGD(lr=-1) on a zeroed buffer produce the dense gradient exactly and
    deterministically (0 - (-1*g) = g), (2) hbAllreduceSumF32 over the peer windows,
    rank-order sum scaled by 1/W, (3) the same fused kernels apply the optimizer with
    ids = 0..rows-1 (one entry per row = a dense apply)."""
    L = _lib.lib()
    W = self.collective.world_size
    m = len(self.local_idx)
    key = (grad.dtype, grad.device)
    if self._dense.get('key') != key:
      total = sum(_weight_of(self.tables[k]).numel() for k in self.local_idx)
      self._dense = {'key': key, 'flat': torch.empty(total, dtype=torch.float32, device=self.device),
                     'rows': [torch.arange(_weight_of(self.tables[k]).shape[0], dtype=torch.int64,
                                           device=self.device) for k in self.local_idx]}
    flat = self._dense['flat']
    flat.zero_()
    views, o = [], 0
    for k in self.local_idx:
      w = _weight_of(self.tables[k])
      views.append(flat[o:o + w.numel()].view(w.shape))
      o += w.numel()
    # (1) densify
    feats = (_lib.hbUpdateFeature * m)()
    for j, k in enumerate(self.local_idx):
      w = _weight_of(self.tables[k])
      feats[j] = _lib.hbUpdateFeature(
          views[j].data_ptr(), None, None, w.shape[0], ids[k].data_ptr(),
          offsets[k].data_ptr() if offsets[k] is not None else None, B, ids[k].numel(),
          grad[:, self.col_offsets[k]:].data_ptr(), grad.stride(0), w.shape[1],
          _lib.COMBINER[self.combiners[k]], 1)
    ws = self._workspace(feats)
    neg = _lib.hbOptimizer(_lib.OPT['sgd'], -1.0, 0.0, 0.0, 0.0, 0, 1)
    if self._sort_done is not None:
      torch.cuda.current_stream().wait_event(self._sort_done)
      self._sort_done = None
      _lib.check(L.hbGroupSparseApply(
          m, feats, _lib.C.byref(neg), _lib.C.c_void_p(ws.data_ptr()), _lib.C.c_size_t(ws.numel()),
          _lib.C.c_void_p(st.data_ptr()), _util.stream_ptr()), 'GroupLookup densify')
    else:
      _lib.check(L.hbGroupLookupBackwardUpdate(
          m, feats, _lib.C.byref(neg), _lib.C.c_void_p(ws.data_ptr()), _lib.C.c_size_t(ws.numel()),
          _lib.C.c_void_p(st.data_ptr()), _util.stream_ptr()), 'GroupLookup densify')
    # (2) all-reduce, mean
    self.collective.allreduce(flat, scale=1.0 / W, out=flat)
    # (3) dense apply
    feats2 = (_lib.hbUpdateFeature * m)()
    for j, k in enumerate(self.local_idx):
      w = _weight_of(self.tables[k])
      slots = self._slots(k, optimizer)
      rows = self._dense['rows'][j]
      feats2[j] = _lib.hbUpdateFeature(
          w.data_ptr(), slots[0].data_ptr() if len(slots) > 0 else None,
          slots[1].data_ptr() if len(slots) > 1 else None, w.shape[0], rows.data_ptr(), None,
          w.shape[0], w.shape[0], views[j].data_ptr(), w.shape[1], w.shape[1], _lib.COMBINER['sum'], 1)
    ws2 = self._workspace(feats2)
    _lib.check(L.hbGroupLookupBackwardUpdate(
        m, feats2, _lib.C.byref(desc), _lib.C.c_void_p(ws2.data_ptr()), _lib.C.c_size_t(ws2.numel()),
        _lib.C.c_void_p(st.data_ptr()), _util.stream_ptr()), 'GroupLookup dense apply')

  def _slots(self, k, optimizer):
    t = self.tables[k]
    if isinstance(t, ShardedEmbeddingWeights):
      return t.ensure_slots(optimizer)
    if not hasattr(self, '_raw_slots'):
      self._raw_slots = {}
    s = self._raw_slots.setdefault(k, [])
    while len(s) < optimizer.num_slots:
      s.append(torch.full_like(t, optimizer.slot_init(len(s))))
    return s

  def slots(self, k):
    t = self.tables[k]
    if isinstance(t, ShardedEmbeddingWeights):
      return t.slots
    return getattr(self, '_raw_slots', {}).get(k, [])
