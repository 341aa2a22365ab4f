"""Cast (wire dtype) and slab-hash cache probe parity on the GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cast_n(hb):
  L = hb._lib.lib()
  xs = [torch.randn(n, device='cuda') for n in (0, 1, 2049, 100000)]
  hs = [torch.empty(x.numel(), dtype=torch.float16, device='cuda') for x in xs]
  st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
  cnt = hb._lib.i64_array([x.numel() for x in xs])
  assert L.hbCastN(4, hb._lib.ptr_array([x.data_ptr() for x in xs]),
                   hb._lib.ptr_array([h.data_ptr() for h in hs]), cnt, 4, 5, st) == 0
  back = [torch.empty_like(x) for x in xs]
  assert L.hbCastN(4, hb._lib.ptr_array([h.data_ptr() for h in hs]),
                   hb._lib.ptr_array([b.data_ptr() for b in back]), cnt, 5, 4, st) == 0
  for x, h, b in zip(xs, hs, back):
    assert torch.equal(h, x.half()) and torch.equal(b, x.half().float())


def _build_cache(oracle, slabs, keys):
  empty = np.iinfo(np.int64).min
  cache = np.full(slabs * 32, empty, np.int64)
  for k in keys:
    s = oracle.murmur3_hash32(int(k)) % slabs
    for _ in range(slabs):
      row = cache[s * 32:(s + 1) * 32]
      free = np.where(row == empty)[0]
      if len(free):
        row[free[0]] = k
        break
      s = (s + 1) % slabs
  return cache


@pytest.mark.parametrize('slabs,nkeys', [(7, 150), (64, 1500), (3, 96)])
def test_cache_lookup(hb, oracle, slabs, nkeys):
  """HbLookup semantics (embedding/lookup_functors.cu.cc:53-149), compared as sets."""
  rng = np.random.RandomState(slabs)
  present = rng.choice(10**6, size=min(nkeys, slabs * 32), replace=False).astype(np.int64)
  cache = _build_cache(oracle, slabs, present)
  q = np.concatenate([rng.choice(present, 700), rng.randint(10**6, 2 * 10**6, 300)]).astype(np.int64)
  rng.shuffle(q)
  n = len(q)
  L = hb._lib.lib()
  d_cache, d_q = torch.from_numpy(cache).cuda(), torch.from_numpy(q).cuda()
  idx = torch.full((n,), -1, dtype=torch.int32, device='cuda')
  pay = torch.full((n,), -1, dtype=torch.int64, device='cuda')
  cnt = torch.zeros(2, dtype=torch.int32, device='cuda')
  rc = L.hbCacheLookup(C.c_void_p(d_cache.data_ptr()), C.c_int64(slabs), C.c_void_p(d_q.data_ptr()), n,
                       C.c_void_p(idx.data_ptr()), C.c_void_p(pay.data_ptr()), C.c_void_p(cnt.data_ptr()),
                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
  assert rc == 0
  nm, nh = cnt.tolist()
  hi, hc, mi, mk = oracle.cache_lookup(cache, q)
  assert nm == len(mi) and nh == len(hi) and nm + nh == n
  idx, pay = idx.cpu().numpy(), pay.cpu().numpy()
  assert set(zip(idx[:nh].tolist(), pay[:nh].tolist())) == set(zip(hi.tolist(), hc.tolist()))
  assert set(zip(idx[n - nm:].tolist(), pay[n - nm:].tolist())) == set(zip(mi.tolist(), mk.tolist()))


def test_lookup_python_surface(hb, oracle):
  """hb.embedding.lookup mirrors the HbLookup op outputs."""
  rng = np.random.RandomState(5)
  slabs = 16
  present = rng.choice(10**5, size=300, replace=False).astype(np.int64)
  cache = _build_cache(oracle, slabs, present)
  q = np.concatenate([rng.choice(present, 200), rng.randint(10**5, 2 * 10**5, 100)]).astype(np.int64)
  hk, hc, mk_idx, mk = hb.embedding.lookup(torch.from_numpy(cache).cuda(), torch.from_numpy(q).cuda())
  hi, hcache, mi, mkeys = oracle.cache_lookup(cache, q)
  assert set(zip(hk.tolist(), hc.tolist())) == set(zip(hi.tolist(), hcache.tolist()))
  assert set(zip(mk_idx.tolist(), mk.tolist())) == set(zip(mi.tolist(), mkeys.tolist()))
  assert np.array_equal(cache[hc.cpu().numpy()], q[hk.cpu().numpy()])
  e = hb.embedding.lookup(torch.from_numpy(cache).cuda(), torch.empty(0, dtype=torch.int64, device='cuda'))
  assert all(t.numel() == 0 for t in e)
  with pytest.raises(ValueError, match='1D'):
    hb.embedding.lookup(torch.zeros(2, 32, dtype=torch.int64, device='cuda'), torch.zeros(1, dtype=torch.int64, device='cuda'))


def test_unique_id_bootstrap_world1(hb):
  """hbGetUniqueId / hbCommCreateFromId (the reference's HbGetNcclId ->
  HbCreateNcclCollective protocol); the multi-rank exchange is test_gpu_multi.py's."""
  uid = hb.distribute.Collective.get_unique_id()
  assert len(uid) == 128 and uid != hb.distribute.Collective.get_unique_id()
  coll = hb.distribute.Collective(0, 1, window_bytes=1 << 20, unique_id=uid)
  coll.barrier()
  torch.cuda.synchronize()
  coll.close()
  with pytest.raises(RuntimeError):
    hb.distribute.Collective(0, 1, window_bytes=1 << 20, unique_id=bytes(128))


def test_h2d_transfer_n(hb):
  """HbH2DTransferN: pinned inputs in one zero-copy kernel, pageable ones by copy; odd
  sizes, empty tensors and mixed dtypes (ids int64, offsets int64, dense float32)."""
  g = torch.Generator().manual_seed(3)
  shapes = [(65536,), (0,), (4097,), (13,), (1,), (300, 7)]
  host = []
  for i, s in enumerate(shapes):
    t = torch.randint(-2**40, 2**40, s, dtype=torch.int64, generator=g) if i % 2 == 0 else torch.randn(s, generator=g)
    host.append(t.pin_memory() if i != 3 else t)   # one pageable tensor
  host.append(torch.arange(1000, dtype=torch.int32).pin_memory()[3:])   # unaligned view (offset 12 B)
  l0 = hb._lib.lib().hbGetLaunchCount()
  out = hb.embedding.h2d_transfer_n(host)
  torch.cuda.synchronize()
  assert hb._lib.lib().hbGetLaunchCount() - l0 == 1   # one kernel for all pinned tensors
  for h, d in zip(host, out):
    assert d.is_cuda and d.dtype == h.dtype and tuple(d.shape) == tuple(h.shape)
    assert torch.equal(d.cpu(), h)
  # reuse of caller-provided outputs, then straight into a lookup
  ids = torch.randint(0, 1000, (512,), dtype=torch.int64).pin_memory()
  d_ids = torch.empty(512, dtype=torch.int64, device='cuda')
  hb.embedding.h2d_transfer_n([ids], [d_ids])
  table = torch.randn(1000, 16, device='cuda')
  got = hb.embedding.embedding_lookup(table, d_ids, check=True)
  assert torch.equal(got.cpu(), table.cpu()[ids])
