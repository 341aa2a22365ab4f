"""hybridbackend_b200 -- B200-native sharded-embedding hot path behind
HybridBackend's operator surface (hb.distribute / hb.embedding).

Only what the path needs lives here:
  csrc/        hand-written sm_100a CUDA kernels + the C-ABI (include/hb_b200.h)
  distribute/  partition_by_modulo[_n], dual modulo, Collective.alltoall[v][_n]
  embedding/   sharding rule, embedding_lookup_sparse, GroupLookup (fused path)
  training/    sparse optimizer descriptors (Adagrad / LazyAdam / SGD)
The CUDA library is mandatory: nothing here computes on the CPU.
"""
from hybridbackend_b200 import _lib
from hybridbackend_b200 import distribute
from hybridbackend_b200 import embedding
from hybridbackend_b200 import training

__version__ = '0.1.0'


def build(force=False, verbose=False):
  return _lib.build(force=force, verbose=verbose)
