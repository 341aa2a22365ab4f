"""torch <-> C-ABI glue (device pointers, current stream, workspaces)."""
import torch

from hybridbackend_b200 import _lib

_TORCH_DT = {torch.int32: 0, torch.int64: 1, torch.float32: 4, torch.float16: 5,
             torch.int8: 6, torch.uint8: 7, torch.float64: 8}
for _name, _code in (('uint32', 2), ('uint64', 3)):
  if hasattr(torch, _name):
    _TORCH_DT[getattr(torch, _name)] = _code


def dtype_code(t):
  if t.dtype not in _TORCH_DT:
    raise TypeError(f'unsupported dtype {t.dtype}')
  return _TORCH_DT[t.dtype]


def require_cuda(t, what):
  if not isinstance(t, torch.Tensor):
    raise TypeError(f'{what}: expected a torch.Tensor, got {type(t)}')
  if not t.is_cuda:
    raise RuntimeError(
        f'{what}: tensor is on {t.device}; hybridbackend_b200 has no CPU path '
        '(the sm_100a CUDA library is the only implementation)')
  if not t.is_contiguous():
    raise ValueError(f'{what}: tensor must be contiguous')
  return t


def stream_ptr():
  return _lib.C.c_void_p(torch.cuda.current_stream().cuda_stream)


_ws_cache = {}


def workspace(nbytes, device, tag='ws'):
  """A cached uint8 scratch buffer of at least nbytes on `device`."""
  key = (tag, device)
  buf = _ws_cache.get(key)
  if buf is None or buf.numel() < nbytes:
    buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    _ws_cache[key] = buf
  return buf


_status = {}


def status_word(device):
  """Sticky device status word of the calling thread (one rank per thread when an
  in-process group of communicators shares a device)."""
  import threading  # pylint: disable=import-outside-toplevel
  key = (torch.device(device), threading.get_ident())
  w = _status.get(key)
  if w is None:
    w = torch.zeros(1, dtype=torch.int32, device=device)
    _status[key] = w
  return w


def check_status(device, clear=True):
  """Synchronising read of the sticky device status word; raises on error bits
  (the reference's CPU gather raises InvalidArgument for out-of-range ids)."""
  w = status_word(device)
  v = int(w.item())
  if clear and v:
    w.zero_()
  if v & _lib.STATUS_ID_OUT_OF_RANGE:
    raise IndexError('embedding id out of range for its table')
  if v & _lib.STATUS_BAD_OFFSETS:
    raise ValueError('bag offsets are not non-decreasing / within nnz')
  if v & _lib.STATUS_PEER_TIMEOUT:
    raise RuntimeError('timed out waiting for a peer rank (a rank crashed or ranks issued '
                       'different collective sequences)')
  if v & _lib.STATUS_WINDOW_OVERFLOW:
    raise RuntimeError('sharded lookup receive window overflow (raise capacity_factor)')
  return v
