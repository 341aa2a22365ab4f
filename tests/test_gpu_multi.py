"""Multi-GPU parity (needs >= 2 GPUs; one process per GPU, bootstrap over gloo): the
cases of tests/rank_cases.py over real IPC-mapped peer windows and NVLink.  The same
cases run on a single GPU in tests/test_gpu_local_group.py."""
import os
import socket
import sys
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


class ProcEnv:
  def __init__(self, rank, world, hb, oracle):
    self.rank, self.world = rank, world
    self.device = torch.device('cuda', rank)
    self.hb, self.oracle = hb, oracle
    self.autograd = True

  def collective(self, window_bytes):
    return self.hb.distribute.Collective(self.rank, self.world, window_bytes=window_bytes,
                                         device=self.device)

  def barrier(self):
    dist.barrier()


def _worker(rank, world, port, case_name, q):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  try:
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('cpu:gloo,cuda:nccl', rank=rank, world_size=world)
    import hybridbackend_b200 as hb
    from oracle import hb_oracle as o
    import rank_cases
    getattr(rank_cases, case_name)(ProcEnv(rank, world, hb, o))
    torch.cuda.synchronize()
    dist.barrier()
    q.put((rank, 'ok'))
  except Exception:  # pylint: disable=broad-except
    q.put((rank, traceback.format_exc()))
  finally:
    try:
      dist.destroy_process_group()
    except Exception:  # pylint: disable=broad-except
      pass


def _spawn(case_name, world=2):
  if torch.cuda.device_count() < world:
    pytest.skip(f'needs {world} GPUs (the same case runs on one GPU in test_gpu_local_group.py)')
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, case_name, q)) for r in range(world)]
  for p in procs:
    p.start()
  import queue as _queue
  res = []
  try:
    for _ in range(world):
      try:
        res.append(q.get(timeout=240 if not res else 60))
      except _queue.Empty:
        res.append((-1, 'a rank never reported (hung in a kernel or crashed hard)'))
        break
  finally:
    for p in procs:
      p.join(timeout=10)
      if p.is_alive():
        p.kill()
  bad = [r for r in res if r[1] != 'ok']
  assert not bad, '\n'.join(f'rank {r}: {m}' for r, m in bad)


def test_alltoallv_golden_2gpu():
  _spawn('alltoallv_golden')


def test_alltoallv_gradients_golden_2gpu():
  _spawn('alltoallv_grads')


def test_alltoallv_vs_oracle_2gpu():
  _spawn('alltoallv_random')


def test_allreduce_2gpu():
  _spawn('allreduce_case')


def test_sharded_group_lookup_2gpu():
  _spawn('sharded_lookup')


def test_sharded_overflow_is_reported_2gpu():
  _spawn('sharded_overflow')


def test_sharded_plan_recreate_2gpu():
  _spawn('sharded_plan_recreate')


def test_bootstrap_from_unique_id_2gpu():
  _spawn('bootstrap_from_unique_id')


@pytest.mark.parametrize('case', ['sharded_lookup', 'sharded_lookup_dim64_hot', 'sharded_many_features',
                                  'sharded_hot_keys_lazy_adam'])
def test_sharded_group_lookup_8gpu(case):
  _spawn(case, world=8)
