"""Collective: the communicator resource and alltoall[v][_n], host side of K2.

Mirrors hybridbackend/tensorflow/distribute/collective.py:37-350:
  Collective.get()                      process-wide communicator (lazy)
  alltoall(value, sizes=None, common_shape=None, topology=Topology.ALL)
`value` may be a list (the packed AlltoallvN form the reference's Pack pass
produces, graph/optimize_collective.cc:114-119).  At world_size == 1 every
collective is the identity (collective.py:166-167, :252-253).  Bootstrap: the
128-byte tokens are all-gathered over torch.distributed (any backend) -- the
reference broadcasts its NCCL id over TF gRPC (collective.py:108-115).
"""
import ctypes as C

import torch

from hybridbackend_b200 import _lib
from hybridbackend_b200 import _util


class _AlltoallFn(torch.autograd.Function):
  """alltoall_fn of collective.py:302-319: the gradient of an equal-split
  alltoall is the alltoall of the upstream gradient."""

  @staticmethod
  def forward(ctx, coll, value):
    ctx.coll = coll
    return coll._alltoall_equal([value.detach()])[0]

  @staticmethod
  def backward(ctx, grad):
    return None, ctx.coll.alltoall_grad(grad)


class _AlltoallvFn(torch.autograd.Function):
  """alltoallv_fn of collective.py:324-348: the gradient of
  alltoall(value, sizes) is the alltoallv of the (dense) upstream gradient with
  the RECEIVED sizes, so every row's gradient returns to the rank it came from."""

  @staticmethod
  def forward(ctx, coll, value, sizes, common_shape):
    outs, osz = coll._alltoallv_n([value.detach()], [sizes],
                                  None if common_shape is None else [common_shape])
    ctx.coll, ctx.common_shape = coll, common_shape
    ctx.save_for_backward(osz[0])
    ctx.mark_non_differentiable(osz[0])
    return outs[0], osz[0]

  @staticmethod
  def backward(ctx, grad, _sizes_grad):
    (osz,) = ctx.saved_tensors
    return None, ctx.coll.alltoall_grad(grad, osz, ctx.common_shape), None, None


class Topology:
  ALL = 0
  INTRA_NODE = 1
  INTER_NODE = 2


class Collective:
  _instance = None

  def __init__(self, rank, world_size, local_size=None, window_bytes=64 << 20,
               device=None, token_allgather=None, _handle=None, unique_id=None):
    self.rank, self.world_size = int(rank), int(world_size)
    self.local_size = int(local_size or world_size)
    self.device = torch.device(device if device is not None else
                               f'cuda:{torch.cuda.current_device()}')
    self._comm = C.c_void_p()
    self._h_sizes = None
    L = _lib.lib()
    if _handle is not None:  # member of an in-process group (local_group)
      self._comm = _handle
      self.window_bytes = int(L.hbCommWindowBytes(self._comm))
      return
    if unique_id is not None:
      # the reference's protocol (distribute/collective.py:108-115): one 128-byte id made
      # by rank 0 (Collective.get_unique_id) and broadcast by the caller
      if len(unique_id) != _lib.TOKEN_BYTES:
        raise ValueError('unique_id must be the 128 bytes Collective.get_unique_id() returned')
      buf = (C.c_ubyte * _lib.TOKEN_BYTES).from_buffer_copy(bytes(unique_id))
      with torch.cuda.device(self.device):
        _lib.check(L.hbCommCreateFromId(buf, self.rank, self.world_size, self.local_size,
                                        C.c_size_t(window_bytes), C.byref(self._comm)), 'Collective(unique_id)')
      self.window_bytes = int(L.hbCommWindowBytes(self._comm))
      return
    token = (C.c_ubyte * _lib.TOKEN_BYTES)()
    with torch.cuda.device(self.device):
      _lib.check(L.hbCommCreate(self.rank, self.world_size, self.local_size,
                                C.c_size_t(window_bytes), C.byref(self._comm), token),
                 'Collective')
      tokens = bytes(token)
      if self.world_size > 1:
        if token_allgather is None:
          token_allgather = _torch_allgather
        all_tokens = token_allgather(tokens, self.world_size)
        if len(all_tokens) != self.world_size * _lib.TOKEN_BYTES:
          raise RuntimeError('token all-gather returned the wrong number of bytes')
        buf = (C.c_ubyte * len(all_tokens)).from_buffer_copy(all_tokens)
        _lib.check(L.hbCommConnect(self._comm, buf), 'Collective.connect')
    self.window_bytes = int(L.hbCommWindowBytes(self._comm))

  @staticmethod
  def get_unique_id():
    """128-byte communicator id, made on one rank and broadcast to the others (the
    counterpart of HbGetNcclId, nccl_get_id.cc:35-70)."""
    buf = (C.c_ubyte * _lib.TOKEN_BYTES)()
    _lib.check(_lib.lib().hbGetUniqueId(buf), 'get_unique_id')
    return bytes(buf)

  @classmethod
  def local_group(cls, world_size, window_bytes=64 << 20, device=None):
    """world_size communicators on ONE device wired to each other in-process
    (hbCommCreateLocalGroup): rank r must then be driven by its own host thread,
    exactly as a process per GPU would drive it.  Lets a single-GPU box run the
    multi-rank kernels (CI), or W logical shards share one GPU."""
    dev = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
    handles = (C.c_void_p * world_size)()
    with torch.cuda.device(dev):
      _lib.check(_lib.lib().hbCommCreateLocalGroup(int(world_size), C.c_size_t(window_bytes), handles),
                 'Collective.local_group')
    return [cls(r, world_size, device=dev, _handle=C.c_void_p(handles[r])) for r in range(world_size)]

  # -- lifecycle ---------------------------------------------------------------
  @classmethod
  def get(cls, **kwargs):
    """Process-wide communicator from torch.distributed's rank/world (or 1)."""
    if cls._instance is None:
      import torch.distributed as dist  # pylint: disable=import-outside-toplevel
      if dist.is_available() and dist.is_initialized():
        cls._instance = cls(dist.get_rank(), dist.get_world_size(), **kwargs)
      else:
        cls._instance = cls(0, 1, **kwargs)
    return cls._instance

  def close(self):
    if self._comm:
      _lib.lib().hbCommDestroy(self._comm)
      self._comm = C.c_void_p()
    if Collective._instance is self:
      Collective._instance = None

  @property
  def handle(self):
    return self._comm

  def _attach_status(self):
    st = _util.status_word(self.device)
    _lib.lib().hbCommSetStatusWord(self._comm, C.c_void_p(st.data_ptr()))
    return st

  def barrier(self):
    with torch.cuda.device(self.device):
      self._attach_status()
      _lib.check(_lib.lib().hbCommBarrier(self._comm, _util.stream_ptr()), 'barrier')

  def allreduce(self, value, scale=1.0, out=None):
    """Sum of `value` (float32) over the ranks, times `scale`; every rank adds the
    contributions in rank order (bit-identical replicas).  The dense-gradient
    path of replicated small tables (training/gradient.py:157-160; scale=1/W is
    the mean of :77-97) -- W x the bytes of a ring all-reduce, sized for those."""
    _util.require_cuda(value, 'allreduce: value')
    if value.dtype != torch.float32:
      raise TypeError('allreduce: float32 only')
    if out is None:
      out = torch.empty_like(value)
    if self.world_size == 1:
      out.copy_(value)
      if scale != 1.0:
        out.mul_(scale)
      return out
    with torch.cuda.device(self.device):
      st = self._attach_status()
      _lib.check(_lib.lib().hbAllreduceSumF32(
          self._comm, C.c_void_p(value.data_ptr()), C.c_void_p(out.data_ptr()),
          C.c_int64(value.numel()), C.c_float(scale), C.c_void_p(st.data_ptr()),
          _util.stream_ptr()), 'allreduce')
    return out

  # -- alltoall ------------------------------------------------------------------
  def _cast_n(self, tensors, to_dtype):
    """fp32 <-> fp16 for N tensors in one launch (hbCastN; the reference's CastN,
    common/cast.cu.cc:84-495, used when comm_wire_dtype is float16)."""
    outs = [torch.empty(t.shape, dtype=to_dtype, device=t.device) for t in tensors]
    if not tensors:
      return outs
    with torch.cuda.device(tensors[0].device):
      _lib.check(_lib.lib().hbCastN(
          len(tensors), _lib.ptr_array([t.data_ptr() for t in tensors]),
          _lib.ptr_array([o.data_ptr() for o in outs]), _lib.i64_array([t.numel() for t in tensors]),
          _util.dtype_code(tensors[0]), _lib.DTYPE['float16' if to_dtype == torch.float16 else 'float32'],
          _util.stream_ptr()), 'cast')
    return outs

  def alltoall(self, value, sizes=None, common_shape=None, topology=Topology.ALL,
               name=None, wire_dtype=None, check=True):
    """Shuffle value partitions across devices (collective.py:271-350).

    sizes=None: equal split of dim 0 across ranks (HbNcclAlltoall); otherwise
    `sizes[r]` rows of `value` go to rank r and (output, output_sizes) is
    returned, output = concat over source ranks of the segments addressed to me.
    wire_dtype=torch.float16 sends float32 payloads as half on the wire (lossy; the
    reference's `comm_wire_dtype` option, collective.py:291-296,
    nccl_alltoallv.cc:57-88)."""
    del name
    self._check = check
    if topology != Topology.ALL:
      raise NotImplementedError('only Topology.ALL is built (single NVSwitch domain); '
                                'INTRA/INTER_NODE belong to the multi-node path')
    single = isinstance(value, torch.Tensor)
    if single and value.requires_grad and wire_dtype is None:
      # differentiable form (collective.py:302-348)
      if sizes is None:
        return _AlltoallFn.apply(self, value)
      return _AlltoallvFn.apply(self, value, sizes, common_shape)
    values = [value] if single else list(value)
    half_wire = (wire_dtype == torch.float16 and self.world_size > 1 and
                 all(v.dtype == torch.float32 for v in values))
    if half_wire:
      values = self._cast_n([_util.require_cuda(v, 'alltoall: value') for v in values], torch.float16)
    if sizes is None:
      outs = self._alltoall_equal(values)
      if half_wire:
        outs = self._cast_n(outs, torch.float32)
      return outs[0] if single else outs
    szs = [sizes] if single else list(sizes)
    shapes = None if common_shape is None else ([common_shape] if single else list(common_shape))
    outs, osz = self._alltoallv_n(values, szs, shapes)
    if half_wire:
      outs = self._cast_n(outs, torch.float32)
    if single:
      return outs[0], osz[0]
    return outs, osz

  def alltoall_grad(self, value_grad, exchanged_sizes=None, common_shape=None):
    """Gradient of alltoall w.r.t. its `value` (collective.py:306-319, :334-348):
    the upstream gradient shuffled back with the RECEIVED sizes.  This is what the
    autograd form calls; exposed for graph builders that wire gradients themselves."""
    if exchanged_sizes is None:
      return self._alltoall_equal([value_grad.contiguous()])[0]
    outs, _ = self._alltoallv_n([value_grad.contiguous()], [exchanged_sizes],
                                None if common_shape is None else [common_shape])
    return outs[0]

  def _alltoall_equal(self, values):
    W = self.world_size
    if W == 1:
      return values
    sizes = []
    for v in values:
      if v.shape[0] % W != 0:
        raise ValueError('alltoall: dim 0 must be divisible by the world size')
      sizes.append(torch.full((W,), v.shape[0] // W, dtype=torch.int32, device=v.device))
    shapes = [tuple(v.shape[1:]) for v in values]
    outs, _ = self._alltoallv_n(values, sizes, shapes)
    return outs

  def _alltoallv_n(self, values, sizes, common_shapes):
    W = self.world_size
    n = len(values)
    if W == 1:  # collective.py:252-253
      return values, sizes
    L = _lib.lib()
    for k in range(n):
      _util.require_cuda(values[k], 'alltoall: value')
      _util.require_cuda(sizes[k], 'alltoall: sizes')
      if sizes[k].dtype != torch.int32 or sizes[k].numel() != W:
        raise TypeError('alltoall: sizes must be int32 [world_size]')
    if common_shapes is None:
      common_shapes = [tuple(v.shape[1:]) for v in values]
    dev = values[0].device
    recv = [torch.empty(W, dtype=torch.int32, device=dev) for _ in range(n)]
    if self._h_sizes is None or self._h_sizes.numel() < n * W:
      self._h_sizes = torch.empty(max(n * W, 256), dtype=torch.int32).pin_memory()
    with torch.cuda.device(dev):
      self._attach_status()
      _lib.check(L.hbAlltoallvNSizes(
          self._comm, n, _lib.ptr_array([s.data_ptr() for s in sizes]),
          _lib.ptr_array([r.data_ptr() for r in recv]), C.c_void_p(self._h_sizes.data_ptr()),
          _util.stream_ptr()), 'alltoallv sizes')
      # the output shape is data dependent: block the host like the reference
      # does (nccl_alltoallv.cc:533)
      torch.cuda.current_stream().synchronize()
      h = self._h_sizes[:n * W].view(n, W)
      outs = []
      common = []
      for k in range(n):
        cs = tuple(int(d) for d in common_shapes[k])
        tot = int(h[k].sum())
        outs.append(torch.empty((tot,) + cs, dtype=values[k].dtype, device=dev))
        c = 1
        for d in cs:
          c *= d
        common.append(c)
      st = _util.status_word(dev)
      _lib.check(L.hbAlltoallvN(
          self._comm, n, _lib.ptr_array([v.data_ptr() for v in values]),
          _lib.i64_array(common), _lib.i32_array([v.element_size() for v in values]),
          _lib.ptr_array([o.data_ptr() for o in outs]), C.c_void_p(st.data_ptr()),
          _util.stream_ptr()), 'alltoallv')
      if getattr(self, '_check', True):
        # the host already blocked for the sizes (as the reference does); one more
        # status read turns a receive-window overflow or a missing peer into an
        # exception instead of uninitialised outputs
        try:
          _util.check_status(dev)
        except RuntimeError as e:
          if 'overflow' in str(e):
            raise RuntimeError(
                f'alltoall: a rank receives more than half of the free window '
                f'({self.window_bytes} B); create the Collective with a larger window_bytes') from e
          raise
    return outs, recv


def _torch_allgather(token_bytes, world_size):
  import torch.distributed as dist  # pylint: disable=import-outside-toplevel
  if not (dist.is_available() and dist.is_initialized()):
    raise RuntimeError('Collective with world_size > 1 needs torch.distributed '
                       'initialised (or a custom token_allgather)')
  gathered = [None] * world_size
  dist.all_gather_object(gathered, token_bytes)
  return b''.join(gathered)
