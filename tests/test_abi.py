"""The C-ABI library loads and exports every symbol include/hb_b200.h declares;
argument validation works without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  src = open(os.path.join(ROOT, 'include', 'hb_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(hb[A-Z]\w*)\s*\(', src)))


def test_header_symbols_exported(hb):
  L = hb._lib.lib()
  declared = _declared_symbols()
  assert len(declared) >= 25
  missing = [s for s in declared if not hasattr(L, s)]
  assert not missing, f'not exported: {missing}'
  assert sorted(hb._lib.SYMBOLS) == declared


def test_build_info(hb):
  L = hb._lib.lib()
  assert b'sm_100a' in L.hbGetBuildInfo()
  assert L.hbGetVersion() >= 100


def test_partition_workspace_query_and_validation(hb):
  L = hb._lib.lib()
  need = C.c_size_t(0)
  lens = (C.c_int32 * 3)(65536, 0, 5)
  assert L.hbPartitionWorkspaceBytes(3, lens, 8, C.byref(need)) == 0
  assert need.value >= (16 + 0 + 1) * 8 * 4   # one status row of 8 bins per 4096-id tile
  bad = (C.c_int32 * 1)(-1)
  assert L.hbPartitionWorkspaceBytes(1, bad, 8, C.byref(need)) != 0
  assert b'negative' in L.hbGetLastErrorString()
  # num_partitions < 1 is rejected before anything touches the device
  ptrs = (C.c_void_p * 1)(None)
  rc = L.hbPartitionByModuloN(1, 1, ptrs, (C.c_int32 * 1)(0), 0, ptrs, ptrs, ptrs, None,
                              C.c_size_t(0), None)
  assert rc == 1 and b'num_partitions' in L.hbGetLastErrorString()
  rc = L.hbPartitionByModuloN(4, 1, ptrs, (C.c_int32 * 1)(0), 2, ptrs, ptrs, ptrs, None,
                              C.c_size_t(0), None)
  assert rc == 1 and b'dtype' in L.hbGetLastErrorString()


def test_lookup_validation(hb):
  L = hb._lib.lib()
  f = (hb._lib.hbLookupFeature * 1)(hb._lib.hbLookupFeature(None, 10, None, None, 0, None, 6, 6, 1, 1, 0))
  assert L.hbGroupLookupForward(1, f, None, None) == 1
  assert b'multiple of 4' in L.hbGetLastErrorString()
  assert L.hbGroupLookupForward(0, f, None, None) == 1


def test_update_workspace_query(hb):
  L = hb._lib.lib()
  feats = (hb._lib.hbUpdateFeature * 2)(
      hb._lib.hbUpdateFeature(None, None, None, 40000000, None, None, 65536, 65536, None, 32, 32, 1, 1),
      hb._lib.hbUpdateFeature(None, None, None, 3, None, None, 65536, 65536, None, 32, 32, 1, 1))
  need = C.c_size_t(0)
  assert L.hbGroupSparseUpdateWorkspaceBytes(2, feats, C.byref(need)) == 0
  assert need.value > 2 * 65536 * 16


def test_cpu_tensors_are_rejected_not_computed(hb):
  """No CPU fallback: host tensors raise instead of silently computing."""
  import torch
  with pytest.raises(RuntimeError, match='no CPU path'):
    hb.distribute.partition_by_modulo(torch.arange(10), 2)
  with pytest.raises(RuntimeError, match='no CPU path'):
    hb.embedding.embedding_lookup_sparse(torch.zeros(4, 4), torch.zeros(2, dtype=torch.int64))


def test_product_does_not_import_oracle():
  """The product package must not route through oracle/ (checked textually)."""
  pkg = os.path.join(ROOT, 'hybridbackend_b200')
  for dirpath, _, files in os.walk(pkg):
    for fn in files:
      if fn.endswith(('.py', '.cu', '.cuh', '.h')):
        txt = open(os.path.join(dirpath, fn)).read()
        assert 'hb_oracle' not in txt.replace('oracle/hb_oracle.c:', ''), fn
        assert 'import oracle' not in txt and 'from oracle' not in txt, fn


def test_struct_layouts_match_header(hb, tmp_path):
  """The ctypes mirrors in hybridbackend_b200/_lib.py must have exactly the layout a C
  compiler gives the structs of include/hb_b200.h (sizeof + every field offset)."""
  import subprocess
  lib = hb._lib
  structs = {'hbLookupFeature': lib.hbLookupFeature, 'hbUpdateFeature': lib.hbUpdateFeature,
             'hbOptimizer': lib.hbOptimizer, 'hbShardedFeature': lib.hbShardedFeature}
  lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hb_b200.h"', 'int main(void) {']
  for name, st in structs.items():
    lines.append(f'  printf("{name} sizeof %zu\\n", sizeof({name}));')
    for fname, _ in st._fields_:
      lines.append(f'  printf("{name} {fname} %zu\\n", offsetof({name}, {fname}));')
  lines += ['  return 0;', '}']
  src = tmp_path / 'layout.c'
  src.write_text('\n'.join(lines))
  exe = tmp_path / 'layout'
  subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
  out = subprocess.check_output([str(exe)], text=True)
  for line in out.strip().splitlines():
    name, field, val = line.split()
    st = structs[name]
    if field == 'sizeof':
      assert C.sizeof(st) == int(val), f'{name}: ctypes {C.sizeof(st)} != C {val}'
    else:
      assert getattr(st, field).offset == int(val), f'{name}.{field}'


def test_header_is_plain_c(tmp_path):
  """The boundary header must compile as C99 with no CUDA / C++ / torch includes."""
  import subprocess
  src = tmp_path / 'h.c'
  src.write_text('#include "hb_b200.h"\nint main(void) { return HB_OK; }\n')
  subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                         str(src), '-o', str(tmp_path / 'h')])
  txt = open(os.path.join(ROOT, 'include', 'hb_b200.h')).read()
  code = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)   # comments cite reference paths
  includes = re.findall(r'#include\s*[<"]([^>"]+)[>"]', code)
  assert sorted(includes) == ['stddef.h', 'stdint.h'], includes
  assert 'torch' not in code and 'Tensor' not in code
