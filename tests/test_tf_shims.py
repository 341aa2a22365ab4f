"""The TensorFlow-1.15 OpKernel shims under hybridbackend_b200/csrc/tf_ops cannot be
built here (no TensorFlow in the image), but they must at least be valid C++ against the
TF-1.15 API they use and against include/hb_b200.h: type-check every shim with
g++ -fsyntax-only against a declaration-level stand-in of that API
(oracle/tf_shim_stub, test infrastructure).  Kernel class templates are instantiated by
the stub's REGISTER_KERNEL_BUILDER, so their bodies are checked too."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = sorted(glob.glob(os.path.join(ROOT, 'hybridbackend_b200', 'csrc', 'tf_ops', '*.cc')))
CUDA_INC = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')


def test_every_op_family_has_a_shim():
  names = {os.path.basename(p) for p in SHIMS}
  assert {'hb_b200_partition_ops.cc', 'hb_b200_collective_ops.cc', 'hb_b200_lookup_ops.cc'} <= names


@pytest.mark.parametrize('shim', SHIMS, ids=[os.path.basename(p) for p in SHIMS])
def test_shim_type_checks(shim):
  gxx = shutil.which('g++')
  if gxx is None or not os.path.exists(os.path.join(CUDA_INC, 'cuda_runtime.h')):
    pytest.skip('needs g++ and the CUDA headers')
  cmd = [gxx, '-std=c++14', '-fsyntax-only', '-Wall', '-Werror', '-DHB_B200_WITH_TENSORFLOW=1',
         '-I' + os.path.join(ROOT, 'oracle', 'tf_shim_stub'), '-I' + os.path.join(ROOT, 'include'),
         '-I' + CUDA_INC, shim]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
  assert r.returncode == 0, r.stderr[-3000:]


def test_shims_register_the_reference_op_names():
  """Kernel registrations use exactly the op names of the reference's REGISTER_OPs."""
  text = ''.join(open(p).read() for p in SHIMS)
  for op in ['HbPartitionByModulo', 'HbPartitionByModuloN', 'HbPartitionByDualModuloStageOne',
             'HbPartitionByDualModuloStageTwoN', 'HbGetNcclId', 'HbCreateNcclCollective',
             'HbNcclAlltoall', 'HbNcclAlltoallN', 'HbNcclAlltoallv', 'HbNcclAlltoallvN', 'HbLookup']:
    assert f'"{op}"' in text, op
