"""Build and load libhb_b200.so (the C-ABI in include/hb_b200.h) through ctypes.

The library is the product: there is NO CPU fallback.  If it is missing or
cannot be loaded, every operator of this package raises.
"""
import ctypes as C
import glob
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, 'csrc')
_SO = os.path.join(_PKG, 'lib', 'libhb_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
              '-std=c++17', '-Xcompiler', '-fPIC']


def sources():
  return sorted(glob.glob(os.path.join(_CSRC, '*.cu')))


def _stale():
  if not os.path.exists(_SO):
    return True
  t = os.path.getmtime(_SO)
  deps = sources() + glob.glob(os.path.join(_CSRC, '*.cuh')) + [
      os.path.join(_ROOT, 'include', 'hb_b200.h')]
  return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
  """nvcc-compile every .cu for sm_100a into hybridbackend_b200/lib/ (in-tree).
  Safe against concurrent callers (N ranks of one torchrun importing the package on a
  fresh box): one builder at a time under a file lock, objects and the library are
  written to temporary names and renamed into place."""
  if not force and not _stale():
    return _SO
  import fcntl  # pylint: disable=import-outside-toplevel
  os.makedirs(os.path.dirname(_SO), exist_ok=True)
  with open(os.path.join(os.path.dirname(_SO), '.build.lock'), 'w') as lock:
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
      if not force and not _stale():   # another process built it while we waited
        return _SO
      return _build_locked(force, verbose)
    finally:
      fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
  nvcc = os.environ.get('NVCC', 'nvcc')
  objdir = os.path.join(_PKG, 'lib', 'obj')
  os.makedirs(objdir, exist_ok=True)
  objs = []
  procs = []
  for src in sources():
    obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
    objs.append(obj)
    hdrs = glob.glob(os.path.join(_CSRC, '*.cuh')) + [os.path.join(_ROOT, 'include', 'hb_b200.h')]
    if (not force and os.path.exists(obj) and
        all(os.path.getmtime(obj) > os.path.getmtime(d) for d in [src] + hdrs)):
      continue
    tmp = obj + f'.tmp{os.getpid()}'
    cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', tmp]
    if verbose:
      print(' '.join(cmd))
    procs.append((cmd, tmp, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
  failed = None
  for cmd, tmp, obj, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0:
      failed = failed or RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), out.decode()))
      if os.path.exists(tmp):
        os.remove(tmp)
    else:
      os.replace(tmp, obj)
  if failed is not None:
    raise failed
  tmp_so = _SO + f'.tmp{os.getpid()}'
  link = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', tmp_so] + objs
  subprocess.check_call(link)
  os.replace(tmp_so, _SO)
  return _SO


class hbLookupFeature(C.Structure):
  _fields_ = [('table', C.c_void_p), ('rows', C.c_int64), ('ids', C.c_void_p),
              ('offsets', C.c_void_p), ('nbags', C.c_int64), ('out', C.c_void_p),
              ('out_stride', C.c_int64), ('dim', C.c_int32), ('combiner', C.c_int32),
              ('id_div', C.c_int64), ('nnz', C.c_int64)]


class hbUpdateFeature(C.Structure):
  _fields_ = [('table', C.c_void_p), ('slot0', C.c_void_p), ('slot1', C.c_void_p),
              ('rows', C.c_int64), ('ids', C.c_void_p), ('offsets', C.c_void_p),
              ('nbags', C.c_int64), ('nnz', C.c_int64), ('grad', C.c_void_p),
              ('grad_stride', C.c_int64), ('dim', C.c_int32), ('combiner', C.c_int32),
              ('id_div', C.c_int64)]


class hbOptimizer(C.Structure):
  _fields_ = [('kind', C.c_int32), ('lr', C.c_float), ('beta1', C.c_float),
              ('beta2', C.c_float), ('eps', C.c_float), ('flags', C.c_int32),
              ('step', C.c_int64)]


class hbShardedFeature(C.Structure):
  _fields_ = [('shard', C.c_void_p), ('slot0', C.c_void_p), ('slot1', C.c_void_p),
              ('shard_rows', C.c_int64), ('ids', C.c_void_p), ('offsets', C.c_void_p),
              ('nbags', C.c_int64), ('nnz', C.c_int64), ('out', C.c_void_p),
              ('out_stride', C.c_int64), ('grad', C.c_void_p), ('grad_stride', C.c_int64),
              ('dim', C.c_int32), ('combiner', C.c_int32)]


HB_OK = 0
DTYPE = {'int32': 0, 'int64': 1, 'uint32': 2, 'uint64': 3, 'float32': 4,
         'float16': 5, 'int8': 6, 'uint8': 7, 'float64': 8}
COMBINER = {'sum': 0, 'mean': 1, 'sqrtn': 2}
OPT = {'sgd': 0, 'adagrad': 1, 'lazy_adam': 2}
STATUS_ID_OUT_OF_RANGE = 1
STATUS_WINDOW_OVERFLOW = 2
STATUS_BAD_OFFSETS = 4
STATUS_PEER_TIMEOUT = 8
OPT_FLAG_FAST_MATH = 1
TOKEN_BYTES = 128

# every symbol include/hb_b200.h declares (tests check the .so exports all)
SYMBOLS = [
    'hbGetLastErrorString', 'hbGetVersion', 'hbGetBuildInfo', 'hbGetLaunchCount',
    'hbProfileEnable', 'hbProfileReset', 'hbProfileGet', 'hbKernelName',
    'hbPartitionWorkspaceBytes', 'hbPartitionByModuloN', 'hbPartitionByDualModuloN',
    'hbGroupLookupForward', 'hbGroupSparseUpdateWorkspaceBytes',
    'hbGroupLookupBackwardUpdate', 'hbGroupSparseSort', 'hbGroupSparseApply', 'hbCastN', 'hbCacheLookup',
    'hbCommCreate', 'hbCommConnect', 'hbGetUniqueId', 'hbCommCreateFromId', 'hbCommCreateLocalGroup', 'hbCommSetStatusWord',
    'hbAllreduceSumF32', 'hbCommDestroy', 'hbCommRank', 'hbCommWorldSize',
    'hbCommWindow', 'hbCommWindowBytes', 'hbCommBarrier',
    'hbAlltoallvNSizes', 'hbAlltoallvN',
    'hbShardedPlanCreate', 'hbShardedPlanDestroy', 'hbShardedPlanWindowBytes',
    'hbShardedLookupForward', 'hbShardedLookupBackwardUpdate',
    'hbGroupLookupForwardHost', 'hbH2DTransferN',
]

_lib = None


def lib():
  """Load the library (building it first if sources are newer).  Raises if the
  CUDA extension cannot be produced: there is no fallback path."""
  global _lib
  if _lib is not None:
    return _lib
  if _stale():
    try:
      build()
    except Exception as e:  # pylint: disable=broad-except
      if not os.path.exists(_SO):
        raise RuntimeError(
            'hybridbackend_b200: libhb_b200.so is missing and could not be built '
            f'({e}); the CUDA extension is required, there is no CPU fallback')
  L = C.CDLL(_SO)
  L.hbGetLastErrorString.restype = C.c_char_p
  L.hbGetBuildInfo.restype = C.c_char_p
  L.hbKernelName.restype = C.c_char_p
  L.hbGetLaunchCount.restype = C.c_int64
  L.hbCommWindow.restype = C.c_void_p
  L.hbCommWindowBytes.restype = C.c_size_t
  L.hbShardedPlanWindowBytes.restype = C.c_size_t
  _lib = L
  return L


def check(rc, what=''):
  if rc != HB_OK:
    msg = lib().hbGetLastErrorString().decode()
    raise RuntimeError(f'hb_b200 {what} failed (status {rc}): {msg}')


def ptr_array(ptrs):
  return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])


def i32_array(vals):
  return (C.c_int32 * len(vals))(*[int(v) for v in vals])


def i64_array(vals):
  return (C.c_int64 * len(vals))(*[int(v) for v in vals])
