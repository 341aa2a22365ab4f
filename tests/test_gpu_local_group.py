"""Multi-rank parity on ONE GPU: every case of tests/rank_cases.py driven by one host
thread per rank through an in-process group of communicators
(hbCommCreateLocalGroup).  The ranks call exactly the entry points, and launch exactly
the kernels, of the one-process-per-GPU deployment (tests/test_gpu_multi.py); only the
peer windows are allocations on the same device instead of IPC-mapped ones, and the
library issues the ranks' kernels phase by phase so that no kernel spins on a flag whose
producer is not yet enqueued (csrc/comm.cuh comm_submit)."""
import os
import sys
import threading
import traceback

import pytest
import torch

os.environ.setdefault('HB_LOCAL_GROUP_TIMEOUT_S', '60')

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rank_cases  # noqa: E402  pylint: disable=wrong-import-position

pytestmark = pytest.mark.gpu


class _Shared:
  def __init__(self, world):
    self.lock = threading.Lock()
    self.groups = {}
    self.barrier = threading.Barrier(world)


class LocalEnv:
  def __init__(self, rank, world, shared, hb, oracle):
    self.rank, self.world = rank, world
    self.device = torch.device('cuda', 0)
    self.hb, self.oracle = hb, oracle
    self.autograd = False
    self._shared = shared
    self._ncoll = 0

  def collective(self, window_bytes):
    with self._shared.lock:
      idx = self._ncoll
      self._ncoll += 1
      if idx not in self._shared.groups:
        self._shared.groups[idx] = self.hb.distribute.Collective.local_group(
            self.world, window_bytes=window_bytes, device=self.device)
    return self._shared.groups[idx][self.rank]

  def barrier(self):
    self._shared.barrier.wait(timeout=120)


def run_ranks(case, world, hb, oracle, own_streams=True):
  shared = _Shared(world)
  errors = [None] * world

  def body(rank):
    try:
      torch.cuda.set_device(0)
      env = LocalEnv(rank, world, shared, hb, oracle)
      if own_streams:
        with torch.cuda.stream(torch.cuda.Stream(device=env.device)):
          case(env)
      else:
        case(env)
    except BaseException:  # pylint: disable=broad-except
      errors[rank] = traceback.format_exc()
      shared.barrier.abort()

  threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
  for t in threads:
    t.start()
  for t in threads:
    t.join(timeout=600)
  hung = [r for r, t in enumerate(threads) if t.is_alive()]
  torch.cuda.synchronize()
  bad = [f'rank {r}:\n{e}' for r, e in enumerate(errors) if e]
  assert not hung, f'ranks {hung} never finished\n' + '\n'.join(bad)
  assert not bad, '\n'.join(bad)


def test_alltoallv_golden(hb, oracle):
  run_ranks(rank_cases.alltoallv_golden, 2, hb, oracle)


def test_alltoallv_golden_shared_stream(hb, oracle):
  run_ranks(rank_cases.alltoallv_golden, 2, hb, oracle, own_streams=False)


def test_alltoallv_gradients_golden(hb, oracle):
  run_ranks(rank_cases.alltoallv_grads, 2, hb, oracle)


@pytest.mark.parametrize('world', [2, 3, 8])
def test_alltoallv_vs_oracle(hb, oracle, world):
  run_ranks(rank_cases.alltoallv_random, world, hb, oracle)


def test_alltoallv_overflow_raises_everywhere(hb, oracle):
  run_ranks(rank_cases.alltoallv_overflow, 2, hb, oracle)


@pytest.mark.parametrize('world', [2, 8])
def test_allreduce(hb, oracle, world):
  run_ranks(rank_cases.allreduce_case, world, hb, oracle)


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_group_lookup(hb, oracle, world):
  run_ranks(rank_cases.sharded_lookup, world, hb, oracle)


def test_sharded_group_lookup_c3_regime(hb, oracle):
  run_ranks(rank_cases.sharded_lookup_dim64_hot, 8, hb, oracle)


def test_sharded_200_features_c4_regime(hb, oracle):
  run_ranks(rank_cases.sharded_many_features, 8, hb, oracle)


def test_sharded_hot_keys_lazy_adam_c5_regime(hb, oracle):
  run_ranks(rank_cases.sharded_hot_keys_lazy_adam, 8, hb, oracle)


@pytest.mark.parametrize('world', [2, 8])
def test_sharded_overflow_is_reported_on_every_rank(hb, oracle, world):
  run_ranks(rank_cases.sharded_overflow, world, hb, oracle)


def test_sharded_plan_recreate(hb, oracle):
  run_ranks(rank_cases.sharded_plan_recreate, 2, hb, oracle)
