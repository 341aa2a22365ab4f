#!/usr/bin/env python
"""bench.py -- pooled-embedding-rows/s of the sharded-embedding hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one synthetic Criteo-1TB-shaped batch
(BASELINE.json configs[1] at N=1: 26 sparse features, the Criteo-Terabyte table
sizes, dim 32, batch 65 536, one id per sample per feature, Zipf(1.05) ids,
Adagrad lr 0.01): fused lookup forward -> [B, 26*D] dense-MLP input, then the
backward of that lookup fused with the sparse Adagrad update.  At N>1
(configs[2], dim 64) tables larger than the batch are row-sharded over the ranks
and go through partition -> NVSwitch push all-to-all -> owner gather -> stitch.
`value` = B*F*N pooled rows / max-over-ranks device time of the K timed steps
(inputs resident in HBM); `e2e` = the same with the ids coming from pinned host
memory and the pooled output copied back to the host inside the timed region,
through the C-ABI host entry (hbGroupLookupForwardHost).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# docs/tutorial/ranking/criteo/data/spec.json:105-355 (Criteo-Terabyte vocabularies)
CRITEO_SIZES = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546,
                403346, 10, 2208, 11938, 155, 4, 976, 14, 39979771, 25641295, 39664984,
                585935, 12972, 108, 36]
NUM_BATCHES = 4  # distinct pre-generated id batches the steps rotate over


def parse_args():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=50)
  p.add_argument('--warmup', type=int, default=5)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--dim', type=int, default=0, help='0: 32 at N=1, 64 at N>1')
  p.add_argument('--batch', type=int, default=65536)
  p.add_argument('--dist', default='zipf', choices=['zipf', 'uniform'])
  p.add_argument('--alpha', type=float, default=1.05)
  p.add_argument('--mode', default='train', choices=['train', 'fwd'])
  p.add_argument('--max-rows', type=int, default=0, help='cap table rows (debug)')
  p.add_argument('--cpu-sample-steps', type=int, default=3)
  p.add_argument('--no-cpu-baseline', action='store_true')
  p.add_argument('--no-e2e', action='store_true')
  p.add_argument('--no-uniform', action='store_true')
  p.add_argument('--capacity-factor', type=float, default=0.0, help='0: world size (always safe)')
  return p.parse_args()


def table_sizes(args):
  s = list(CRITEO_SIZES)
  if args.max_rows:
    s = [min(n, args.max_rows) for n in s]
  return s


def gen_ids_numpy(rng, n, vocab, dist, alpha, salt=0):
  """Bounded power-law ranks by inverse CDF, scrambled over the vocabulary.  `salt`
  (the feature index) shifts the scramble, so the hottest id of every feature is a
  different row -- and, row-sharded, a different owner rank (salt 0 would make rank 0
  the owner of the hottest row of EVERY table)."""
  if dist == 'uniform' or vocab <= 2:
    return rng.randint(0, vocab, n).astype(np.int64)
  u = rng.random_sample(n)
  a = 1.0 - alpha
  rank = np.floor(((vocab ** a - 1.0) * u + 1.0) ** (1.0 / a)).astype(np.int64)
  rank = np.clip(rank, 1, vocab) - 1
  return (rank * 2654435761 + (salt * 0x9E3779B1) % vocab) % vocab


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.proc, self.lines = index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
           '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    self.t.join(timeout=2)
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for l in self.lines:
      f = [x.strip() for x in l.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2]))
      except ValueError:
        continue
      for nme, v in zip(names, f[5:9]):
        if v.lower().startswith('active'):
          reasons.add(nme)
    return {'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': max(mx) if mx else None, 'samples': len(sm),
            'reasons': sorted(reasons)}


# --------------------------------------------------------------------------------
# CPU arm: the oracle (reference-semantics restatement) on the host cores
# --------------------------------------------------------------------------------
def cpu_arm(args, sizes, dim, steps, warmup, rank0_only_note=''):
  """Times oracle/ (TF-1.15 embedding_lookup_sparse: unique->gather->segment
  reduce; then dedup + SparseApplyAdagrad) on the host cores; features in
  parallel threads (ctypes releases the GIL).  Returns (rows_per_s, info)."""
  from concurrent.futures import ThreadPoolExecutor
  import psutil
  from oracle import hb_oracle as o
  o.lib()
  F, B = len(sizes), args.batch
  cores = os.cpu_count() or 1
  avail = psutil.virtual_memory().available
  need = sum(sizes) * dim * 4 * (2 if args.mode == 'train' else 1)
  scale = 1.0
  if need > 0.6 * avail:
    scale = 0.6 * avail / need
  csizes = [max(1, int(n * scale)) for n in sizes]
  rng = np.random.RandomState(1234)
  tables = [np.zeros((n, dim), np.float32) for n in csizes]       # lazily paged
  accs = [np.zeros((n, dim), np.float32) for n in csizes] if args.mode == 'train' else None
  nb = min(NUM_BATCHES, 2)
  batches = [[gen_ids_numpy(rng, B, n, args.dist, args.alpha, salt=k) for k, n in enumerate(csizes)] for _ in range(nb)]
  offsets = np.arange(B + 1, dtype=np.int64)
  grad = rng.randn(B, F * dim).astype(np.float32)
  out = np.empty((B, F * dim), np.float32)
  for k in range(F):  # touch the rows the batches use so page faults are not timed
    for b in batches:
      u = np.unique(b[k])
      tables[k][u] = 1e-3
      if accs is not None:
        accs[k][u] = 0.1

  # Three work decompositions; the fastest one on this box is the one reported.
  #  "pthreads": C work queue over (feature, bag-chunk) and (feature, row-part) tasks
  #  "feature":  one task per feature doing forward then backward (26 threads)
  #  "split":    forward split into feature x bag-chunk tasks, backward of a big
  #              table split by id % parts (disjoint row sets: the per-part dedup +
  #              Adagrad applies are independent and equal the unsplit update)
  fwd_chunks = max(1, min(8, cores // F))
  bounds = np.linspace(0, B, fwd_chunks + 1).astype(np.int64)

  def one_feature(k, ids):
    o.embedding_lookup_sparse(tables[k], ids, offsets, 'mean', out=out[:, k * dim:(k + 1) * dim])
    if accs is not None:
      o.sparse_apply_adagrad(tables[k], accs[k], ids, np.ascontiguousarray(grad[:, k * dim:(k + 1) * dim]), 0.01)

  def fwd_task(k, c, ids):
    s, e = int(bounds[c]), int(bounds[c + 1])
    o.embedding_lookup_sparse(tables[k], ids[s:e], offsets[:e - s + 1], 'mean',
                              out=out[s:e, k * dim:(k + 1) * dim])

  def parts_of(k):
    return max(1, min(8, cores // F)) if csizes[k] > 1000000 else 1

  def bwd_task(k, part, parts, ids):
    if parts == 1:
      sel_ids, g = ids, grad[:, k * dim:(k + 1) * dim]
    else:
      sel = np.nonzero(ids % parts == part)[0]
      sel_ids, g = ids[sel], grad[sel, k * dim:(k + 1) * dim]
    # mean combiner with one id per bag: row gradient == bag gradient
    o.sparse_apply_adagrad(tables[k], accs[k], sel_ids, np.ascontiguousarray(g), 0.01)

  with ThreadPoolExecutor(cores) as ex:
    def step_feature(i):
      b = batches[i % nb]
      list(ex.map(lambda k: one_feature(k, b[k]), range(F)))

    def step_split(i):
      b = batches[i % nb]
      for f in [ex.submit(fwd_task, k, c, b[k]) for k in range(F) for c in range(fwd_chunks)]:
        f.result()
      if accs is not None:
        for f in [ex.submit(bwd_task, k, p, parts_of(k), b[k]) for k in range(F)
                  for p in range(parts_of(k))]:
          f.result()

    def step_pthreads(i):
      # the same per-feature oracle functions driven by a C pthread work queue
      # (oracle/hb_oracle_mt.c): no Python in the loop
      bt = batches[i % nb]
      o.mt_step(tables, accs, [bt[k] for k in range(F)], grad, out, 0.01, cores,
                fwd_chunk=8192, parts=[parts_of(k) for k in range(F)])

    # time ALL decompositions for the full step count and report the fastest one
    trial = {}
    for name, fn in (('feature', step_feature), ('split', step_split), ('pthreads', step_pthreads)):
      for i in range(max(1, warmup)):
        fn(i)
      t0 = time.perf_counter()
      for i in range(steps):
        fn(i)
      trial[name] = (time.perf_counter() - t0) / steps
    mode = min(trial, key=trial.get)
    workers = cores if mode == 'pthreads' else min(cores, F if mode == 'feature' else F * fwd_chunks)
    dt = trial[mode] * steps
  value = B * F * steps / dt
  info = {'value': value, 'unit': 'pooled-embedding-rows/s', 'cores': workers,
          'host_cores': cores,
          'kind': 'port',
          'sample': (f'{steps} steps x (26 feats x {B} ids) of the same workload, '
                     f'{"fwd+bwd+Adagrad" if args.mode == "train" else "fwd"}, tables '
                     f'{"full size" if scale == 1.0 else f"scaled x{scale:.2f} to fit host RAM"}, '
                     f'{workers} threads (decomposition "{mode}", the fastest of three thread decompositions, s/step: {trial}); oracle/hb_oracle.c port of the '
                     'TF-1.15 CPU semantics (the tf115 wheel cannot run here)' + rank0_only_note),
          'ms_per_step': dt / steps * 1e3}
  return value, info


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  dim = args.dim or (32 if args.gpus == 1 else 64)
  sizes = table_sizes(args)
  steps = max(1, min(args.steps, 5))
  warmup = max(1, min(args.warmup, 2))
  value, info = cpu_arm(args, sizes, dim, steps, warmup)
  line = {
      'impl': 'reference', 'metric': 'pooled-embedding-rows/sec', 'value': value,
      'unit': 'pooled-embedding-rows/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
      'ms_per_step': info['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': dict(workload_config(args, dim, sizes), reference_steps_cap=5, reference_warmup_cap=2),
      'cpu_baseline': info,
      'e2e': {'value': value, 'unit': 'pooled-embedding-rows/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
      'note': ('reference arm = CPU restatement of the reference path (oracle port; TF-1.15 '
               'cannot be installed here), bounded to %d steps' % steps)}
  print(json.dumps(line))


def workload_config(args, dim, sizes):
  return {'workload': ('criteo-1tb-synthetic: 26 sparse feats, table sizes %d..%d rows, dim %d, '
                       'batch %d per rank, 1 id/sample/feature, %s ids, %s' %
                       (min(sizes), max(sizes), dim, args.batch,
                        f'zipf({args.alpha})' if args.dist == 'zipf' else 'uniform',
                        'fwd + bwd + sparse Adagrad(lr=0.01)' if args.mode == 'train' else 'fwd only')),
          'global_batch': args.batch * args.gpus, 'features': len(sizes), 'dim': dim,
          'parallelism': 'single-gpu' if args.gpus == 1 else f'row-sharded tables x{args.gpus} (id % W), data-parallel batch',
          'cache': ('inputs larger than L2: %.1f GB of tables, steps rotate over %d id batches' %
                    (sum(sizes) * dim * 4 / 1e9, NUM_BATCHES))}


# --------------------------------------------------------------------------------
def run_ours(args):
  import torch
  import torch.distributed as dist
  import hybridbackend_b200 as hb
  from hybridbackend_b200 import _lib
  L = _lib.lib()
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit('launch N>1 with torch.distributed.run (one rank per GPU)')
  torch.cuda.set_device(local_rank)
  dev = torch.device(f'cuda:{local_rank}')
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  dim = args.dim or (32 if world == 1 else 64)
  sizes = table_sizes(args)
  F, B = len(sizes), args.batch
  if args.capacity_factor <= 0:
    args.capacity_factor = float(world)

  # ---- tables (row-sharded where the reference shards them) ----------------------
  coll = None
  if world > 1:
    from hybridbackend_b200.embedding.sharded import group_window_bytes
    wbytes = group_window_bytes(world, sizes, [dim] * F, [B] * F, args.capacity_factor)
    coll = hb.distribute.Collective(rank, world, window_bytes=wbytes, device=dev)
  g = torch.Generator(device=dev).manual_seed(1234 + rank)
  tables = []
  for k, n in enumerate(sizes):
    t = hb.embedding.ShardedEmbeddingWeights(f"emb{k}", n, dim, rank, world, batch_size=-1, device=dev)
    t.weight.uniform_(-1e-3, 1e-3, generator=g)
    tables.append(t)
  gl = hb.embedding.GroupLookup(tables, ['mean'] * F, collective=coll, max_nnz=[B] * F,
                                capacity_factor=args.capacity_factor)
  opt = hb.training.Adagrad(0.01)

  # ---- inputs: NUM_BATCHES id batches, device + pinned host copies ---------------
  rng = np.random.RandomState(1234 + rank)
  h_batches = []
  for _ in range(NUM_BATCHES):
    blk = np.stack([gen_ids_numpy(rng, B, n, args.dist, args.alpha, salt=k) for k, n in enumerate(sizes)])  # [F, B]
    h_batches.append(torch.from_numpy(blk).pin_memory())
  d_batches = [hb_.to(dev) for hb_ in h_batches]
  grad = torch.randn(B, F * dim, device=dev, generator=g)
  out = torch.empty(B, F * dim, device=dev)
  # end-to-end leg: two (device out, pinned host out, id staging) sets, so the D2H of step t
  # (on a copy stream) overlaps the backward of step t and the H2D + forward of step t+1
  outs = [out, torch.empty(B, F * dim, device=dev)]
  h_outs = [torch.empty(B, F * dim).pin_memory() for _ in range(2)]
  d_stages = [torch.empty(F, B, dtype=torch.int64, device=dev) for _ in range(2)]
  d2h = torch.cuda.Stream(device=dev)
  h_metric = torch.zeros(1, dtype=torch.int32).pin_memory()

  def step(i, host=False):
    if host == 'device_out':
      # ids from pinned host memory, pooled output stays on the device (its consumer is the
      # on-device MLP); the step's D2H is the 4-byte status word
      d_stages[i & 1].copy_(h_batches[i % NUM_BATCHES], non_blocking=True)
      gl.forward([d_stages[i & 1][k] for k in range(F)], out=outs[i & 1])
    elif host:
      gl.forward_host(h_batches[i % NUM_BATCHES], d_stages[i & 1], outs[i & 1], h_outs[i & 1], d2h_stream=d2h)
    else:
      db = d_batches[i % NUM_BATCHES]
      gl.forward([db[k] for k in range(F)], out=out)
    if args.mode == 'train':
      gl.backward_update(grad, opt)
    if host == 'device_out':
      h_metric.copy_(hb._util.status_word(dev), non_blocking=True)

  def timed(steps, host=False):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for i in range(steps):
      step(i, host)
    timed.host_ms = (time.perf_counter() - t_host) * 1e3 / steps  # host enqueue time per step
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
      dist.barrier()
    return ms

  W_, K = max(args.warmup, 3), args.steps
  # nvidia-smi needs ~1 s to start sampling: start it before the warm-up so that
  # samples exist for the (short) timed region; rank 0 samples its own GPU
  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()
    time.sleep(1.0)
  for i in range(W_):
    step(i)
  hb._util.check_status(dev)
  if rank == 0:
    torch.cuda.synchronize()
    clocks.lines.clear()   # keep only samples taken from here on (timed region under load)
  l0 = L.hbGetLaunchCount()
  ms = timed(K)
  host_ms = timed.host_ms
  launches = L.hbGetLaunchCount() - l0
  # the timed region is only tens of ms: keep the GPUs under the same load for
  # ~0.5 s more (same step count on every rank) so the 100 ms sampler sees it
  extra = min(3000, max(10, int(0.5e3 * K / ms)))
  for i in range(extra):
    step(i)
  torch.cuda.synchronize()
  clk = clocks.stop() if rank == 0 else None
  if world > 1:
    dist.barrier()
  hb._util.check_status(dev)
  value = B * F * world * K / (ms * 1e-3)

  # ---- per-kernel CUDA-event pass (same K steps) for the roofline ----------------
  # (the backward's id sort normally overlaps the forward on a side stream; for
  # clean per-kernel durations this pass runs everything on one stream)
  gl.overlap_backward_sort = False
  step(0)
  L.hbProfileReset(); L.hbProfileEnable(1)
  timed(K)
  L.hbProfileEnable(0)
  gl.overlap_backward_sort = True
  kern = {}
  for kid in range(1, 32):
    tms, n = C.c_double(0), C.c_int64(0)
    L.hbProfileGet(kid, C.byref(tms), C.byref(n))
    if n.value:
      kern[L.hbKernelName(kid).decode()] = {'ms_total': tms.value, 'launches': n.value,
                                            'ms_avg': tms.value / n.value}
  L.hbProfileReset()
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
  except (OSError, ValueError):
    pass
  peak = float(peaks.get('hbm_gbs', 6650.0))
  peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
  local_feats = [k for k in range(F) if k in gl.local_idx]
  shard_feats = [k for k in range(F) if k in gl.sharded_idx]
  alg = {}
  bound = {}
  extra = {}
  kShortMax = 16  # csrc/sparse_update.cu: longer runs go to the hot-row kernel

  def run_stats(idsets):
    """entries / unique rows of the short-run and of the hot-row kernel, mean per batch"""
    se = su = le = lu = 0.0
    for ids in idsets:
      _, c = np.unique(ids, return_counts=True)
      lg = c > kShortMax
      se += c[~lg].sum(); su += (~lg).sum(); le += c[lg].sum(); lu += lg.sum()
    m = float(len(idsets))
    return se / m, su / m, le / m, lu / m

  if world == 1:
    hb_np = [hb_.numpy() for hb_ in h_batches]
    # forward gather+pool: per pooled row L*(8+4D) read + 4D written, L=1
    alg['lookup_fwd'] = F * B * (8 + 4 * dim + 4 * dim)
    se = su = le = lu = 0.0
    npass = 0.0
    for k in range(F):
      a_, b_, c_, d_ = run_stats([blk[k] for blk in hb_np])
      se += a_; su += b_; le += c_; lu += d_
      npass += 1 if sizes[k] + 2 <= 512 else (2 if sizes[k] + 2 <= (1 << 18) else (3 if sizes[k] + 2 <= (1 << 27) else 4))
    # short-run kernel: per entry 4 B (bag) + 4D (gradient row); per unique row 24 B of run
    # arrays + 16D (table + accumulator, read + write)
    alg['sparse_update'] = se * (4 + 4 * dim) + su * (24 + 16 * dim)
    alg['sparse_update_long'] = le * (4 + 4 * dim) + lu * (28 + 16 * dim)
    # sort + runs: 8 B id in, then (key, value) 8 B out + 8 B in per pass, runs 8 B in + 28 B per unique
    alg['sort_pass'] = B * (npass * 16 + F * 8) + (su + lu) * 28
    extra['unique_rows_per_step'] = su + lu
    extra['hot_row_entries_per_step'] = le
    step_alg = alg['lookup_fwd'] + (alg['sparse_update'] + alg['sparse_update_long'] + alg['sort_pass'] if args.mode == 'train' else 0)
  else:
    # regenerate every rank's batches (same seeds): unique ids per (requester, owner)
    gens = [np.random.RandomState(1234 + r) for r in range(world)]
    recv_u = np.zeros(world)          # unique rows asked of owner r (sum over requesters, features)
    sent_u = 0.0                      # unique rows this rank asks for
    for _ in range(NUM_BATCHES):
      blks = [np.stack([gen_ids_numpy(gr, B, n, args.dist, args.alpha, salt=k) for k, n in enumerate(sizes)]) for gr in gens]
      for k in shard_feats:
        for q, blk in enumerate(blks):
          u = np.unique(blk[k])
          cnt = np.bincount(u % world, minlength=world)
          recv_u += cnt / NUM_BATCHES
          if q == rank:
            sent_u += len(u) / NUM_BATCHES
    nsh = len(shard_feats)
    # NVLink-bound: unique rows stored into the requesters' windows, (W-1)/W of them remote
    alg['sharded_owner_gather'] = recv_u[rank] * 4 * dim * (world - 1) / world
    bound['sharded_owner_gather'] = 'nvlink'
    # stitch/pool: 4 B inverse index + 4D row read (window, mostly L2) + 4D written per pooled row
    alg['sharded_stitch_pool'] = nsh * B * (4 + 8 * dim)
    extra['owner_load_unique_rows'] = {'max': float(recv_u.max()), 'mean': float(recv_u.mean()),
                                       'max_over_mean': float(recv_u.max() / max(recv_u.mean(), 1.0))}
    extra['dedup'] = {'ids_per_rank_per_step': nsh * B, 'unique_ids_sent_per_rank_per_step': sent_u,
                      'wire_reduction': nsh * B / max(sent_u, 1.0)}
    # one direction, per rank and step: unique ids out (4 B) + unique rows in + unique gradient sums out
    nv = sent_u * (world - 1) / world * (4 + 4 * dim)
    extra['nvlink_bytes_per_step_per_rank_one_direction'] = nv
    step_alg = None
  NVLINK_PEAK = 770.0  # GB/s per direction per GPU, measured peer copy (B200_PROFILING.md)
  traffic = {}
  try:
    traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
  except (OSError, ValueError):
    pass

  def roofs_of(kern_):
    r_ = []
    for name, a in alg.items():
      if name in kern_ and a > 0:
        ach = a / (kern_[name]['ms_avg'] * 1e-3) / 1e9
        bd = bound.get(name, 'hbm')
        pk = peak if bd == 'hbm' else NVLINK_PEAK
        tr = traffic.get(name) if world == 1 and args.dist == 'zipf' and dim == 32 else None
        r_.append({'kernel': name, 'bound': bd, 'achieved': ach, 'peak': pk, 'unit': 'GB/s',
                   'frac': ach / pk, 'traffic': tr,
                   'traffic_source': 'profiles/traffic.json (ncu --set full, dram bytes read + written per launch, round 2)' if tr else None,
                   'alg_bytes_per_launch': a, 'ms_avg': kern_[name]['ms_avg'],
                   'peak_source': peak_src if bd == 'hbm' else
                   'measured peer copy 770 GB/s/direction (B200_PROFILING.md); kernel time includes the wait for the peers'})
    r_.sort(key=lambda r: -r['ms_avg'])
    return r_
  roofs = roofs_of(kern)
  # the dominant kernel = the longest one on the critical path of the TIMED step: the id
  # sort + run detection is enqueued on a side stream at forward time (N=1) / hoisted under
  # the stitch (owner side, N>1), so it is listed in roofline_all but not picked here
  for r in roofs:
    r['on_critical_path'] = r['kernel'] != 'sort_pass'
  crit = [r for r in roofs if r['on_critical_path']]
  roofline = crit[0] if crit else (roofs[0] if roofs else None)

  # ---- second data point: uniform ids (no hot rows, no L2 help from duplicates) ----
  if world == 1 and args.dist == 'zipf' and not args.no_uniform:
    saved = d_batches
    rngu = np.random.RandomState(4321)
    d_batches = [torch.from_numpy(np.stack([gen_ids_numpy(rngu, B, n, 'uniform', 0.0) for n in sizes])).to(dev)
                 for _ in range(NUM_BATCHES)]
    for i in range(3):
      step(i)
    ms_u = timed(K)
    gl.overlap_backward_sort = False
    L.hbProfileReset(); L.hbProfileEnable(1)
    timed(K)
    L.hbProfileEnable(0)
    gl.overlap_backward_sort = True
    kern_u = {}
    for kid in range(1, 32):
      tms, n_ = C.c_double(0), C.c_int64(0)
      L.hbProfileGet(kid, C.byref(tms), C.byref(n_))
      if n_.value:
        kern_u[L.hbKernelName(kid).decode()] = {'ms_avg': tms.value / n_.value}
    L.hbProfileReset()
    fwd_u = kern_u.get('lookup_fwd', {}).get('ms_avg')
    extra['uniform'] = {'ms_per_step': ms_u / K, 'value': B * F * K / (ms_u * 1e-3),
                        'lookup_fwd_ms': fwd_u,
                        'lookup_fwd_roofline': ({'achieved': alg['lookup_fwd'] / (fwd_u * 1e-3) / 1e9, 'peak': peak,
                                                 'frac': alg['lookup_fwd'] / (fwd_u * 1e-3) / 1e9 / peak,
                                                 'note': 'uniform ids over the full vocabularies: rows of the 8 large tables come from DRAM; '
                                                         'a pure random 128-B row read sustains 4.3 TB/s on this GPU (tools/gather_peak.cu)'}
                                                if fwd_u else None),
                        'kernels_ms': {k_: v_['ms_avg'] for k_, v_ in kern_u.items()}}
    d_batches = saved

  # ---- end-to-end through the host-buffer C-ABI entry ------------------------------
  e2e = None
  e2e_variants = {}
  if not args.no_e2e:
    for i in range(3):
      step(i, host=True)
    ms_e = timed(K, host=True)
    e2e = {'value': B * F * world * K / (ms_e * 1e-3), 'unit': 'pooled-embedding-rows/s',
           'h2d_bytes_per_step': F * B * 8, 'd2h_bytes_per_step': B * F * dim * 4,
           'ms_per_step': ms_e / K,
           'note': 'ids from pinned host memory (hbGroupLookupForwardHost), pooled [B, F*D] output copied back to '
                   'pinned host on a copy stream overlapping the backward and the next step (PCIe-bound: '
                   '%.0f MB D2H per step); upstream gradient stays on the device (it comes from the dense MLP)'
                   % (B * F * dim * 4 / 1e6)}
    for i in range(3):
      step(i, host='device_out')
    ms_d = timed(K, host='device_out')
    e2e_variants['device_out'] = {
        'value': B * F * world * K / (ms_d * 1e-3), 'unit': 'pooled-embedding-rows/s',
        'h2d_bytes_per_step': F * B * 8, 'd2h_bytes_per_step': 4, 'ms_per_step': ms_d / K,
        'note': 'same, but the pooled output stays in HBM for the on-device MLP (what the reference graph '
                'does: the op output is a GPU tensor); D2H = the 4-byte status word'}

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    _, cpu = cpu_arm(args, sizes, dim, args.cpu_sample_steps, 1)

  if rank == 0:
    line = {
        'metric': 'pooled-embedding-rows/sec', 'value': value, 'unit': 'pooled-embedding-rows/s',
        'n_gpus': world, 'steps': K, 'warmup': W_, 'ms_per_step': ms / K, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, dim, sizes), 'clocks': clk, 'e2e': e2e,
        'e2e_variants': e2e_variants,
        'gpu_launches': int(launches), 'host_enqueue_ms_per_step': host_ms, 'roofline': roofline, 'roofline_all': roofs,
        'kernels': kern, 'cpu_baseline': cpu, 'extra': extra,
        'roofline_note': 'per-kernel CUDA events recorded by the library on the launching stream in '
                         'a second pass of the same K steps (hbProfileEnable), single stream (no fwd/sort overlap); '
                         '`roofline` = the longest kernel on the critical path of the timed step (the sort kernel runs '
                         'on a side stream under the forward; every kernel incl. the sort is in roofline_all)',
        'hbm_roofline_rows_per_s_per_gpu': (peak * 1e9 / (step_alg / (B * F)) if step_alg else None),
        'hbm_roofline_note': ('B*F pooled rows / (algorithmic bytes of one step on THIS workload -- duplicates read once per '
                              'entry, table rows once per unique row -- / measured HBM peak); whole step achieves %.2f of it'
                              % (step_alg / (ms / K * 1e-3) / 1e9 / peak)) if step_alg else None,
    }
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  args = parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
