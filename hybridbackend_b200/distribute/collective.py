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


class Topology:
  ALL = 0
  INTRA_NODE = 1
  INTER_NODE = 2


class Collective:
  _instance = None

  def __init__(self, rank, world_size, local_size=None, window_bytes=64 << 20,
               device=None, token_allgather=None):
    self.rank, self.world_size = int(rank), int(world_size)
    self.local_size = int(local_size or world_size)
    self.device = torch.device(device if device is not None else
                               f'cuda:{torch.cuda.current_device()}')
    self._comm = C.c_void_p()
    token = (C.c_ubyte * _lib.TOKEN_BYTES)()
    L = _lib.lib()
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
    self._h_sizes = None

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

  def barrier(self):
    with torch.cuda.device(self.device):
      _lib.check(_lib.lib().hbCommBarrier(self._comm, _util.stream_ptr()), 'barrier')

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
               name=None, wire_dtype=None):
    """Shuffle value partitions across devices (collective.py:271-350).

    sizes=None: equal split of dim 0 across ranks (HbNcclAlltoall); otherwise
    `sizes[r]` rows of `value` go to rank r and (output, output_sizes) is
    returned, output = concat over source ranks of the segments addressed to me.
    wire_dtype=torch.float16 sends float32 payloads as half on the wire (lossy; the
    reference's `comm_wire_dtype` option, collective.py:291-296,
    nccl_alltoallv.cc:57-88)."""
    del name
    if topology != Topology.ALL:
      raise NotImplementedError('only Topology.ALL is built (single NVSwitch domain); '
                                'INTRA/INTER_NODE belong to the multi-node path')
    single = isinstance(value, torch.Tensor)
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
    return outs, recv


def _torch_allgather(token_bytes, world_size):
  import torch.distributed as dist  # pylint: disable=import-outside-toplevel
  if not (dist.is_available() and dist.is_initialized()):
    raise RuntimeError('Collective with world_size > 1 needs torch.distributed '
                       'initialised (or a custom token_allgather)')
  gathered = [None] * world_size
  dist.all_gather_object(gathered, token_bytes)
  return b''.join(gathered)
