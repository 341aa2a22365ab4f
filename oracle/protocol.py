"""Executable numpy specification of the fused sharded exchange (test infrastructure).

`fused_forward` walks the exact buffers and offset formulas that
hybridbackend_b200/csrc/sharded.cu derives on the device from the all-gathered
W x W size matrix (ShFeatMeta: send_off, remote_base, recv_base, src_bucket_off),
so the offset algebra can be checked on a CPU against the plain recipe of
hybridbackend/tensorflow/embedding/sharding.py:171-203 (oracle/hb_oracle.c).
`fused_forward_dedup` is the round-2 variant (requester-side dedup: unique ids on
the wire) specified the same way.
"""
import numpy as np

from oracle import hb_oracle as o


def _meta(S, me):
  """Offsets rank `me` derives from S[q][r] = ids rank q sends to owner r."""
  W = S.shape[0]
  send_off = np.concatenate([[0], np.cumsum(S[me])])                 # my bucket starts
  remote_base = np.array([S[:me, r].sum() for r in range(W)])        # my segment in owner r's window
  recv_base = np.concatenate([[0], np.cumsum(S[:, me])])             # as owner: source q's segment
  src_bucket_off = np.array([S[q, :me].sum() for q in range(W)])     # bucket `me` start at requester q
  return send_off, remote_base, recv_base, src_bucket_off


def fused_forward(shards, bucket_size, ids_per_rank):
  """Returns per rank the [n, dim] rows in input order, computed with the fused
  protocol's windows (ids_in, rows_in) and offsets."""
  W = len(shards)
  dim = shards[0].shape[1]
  parts = [o.partition_by_modulo(np.asarray(i, np.int64), W) for i in ids_per_rank]
  S = np.stack([p[1] for p in parts]).astype(np.int64)               # [q][r]
  meta = [_meta(S, r) for r in range(W)]
  # push_ids: requester q writes bucket r into owner r's ids_in at remote_base
  ids_in = [np.full(int(S[:, r].sum()), -1, np.int64) for r in range(W)]
  for q in range(W):
    send_off, remote_base, _, _ = meta[q]
    for r in range(W):
      seg = parts[q][0][send_off[r]:send_off[r + 1]]
      ids_in[r][remote_base[r]:remote_base[r] + len(seg)] = seg
  assert all((a >= 0).all() for a in ids_in)
  # owner_gather: owner r writes row of received position p into requester q's
  # rows_in at src_bucket_off[q] + (p - recv_base[q])
  rows_in = [np.full((len(ids_per_rank[q]), dim), np.nan, np.float32) for q in range(W)]
  for r in range(W):
    _, _, recv_base, src_bucket_off = meta[r]
    for p, gid in enumerate(ids_in[r]):
      q = int(np.searchsorted(recv_base, p, side='right') - 1)
      rows_in[q][src_bucket_off[q] + (p - recv_base[q])] = shards[r][gid // W]
  # stitch: out[i] = rows_in[idx[i]]
  return [rows_in[q][parts[q][2]] for q in range(W)], ids_in, meta


def fused_forward_dedup(shards, bucket_size, ids_per_rank):
  """Requester-side dedup: only unique ids (sorted) travel; rows come back once per
  unique id and are expanded by the inverse map during the stitch."""
  W = len(shards)
  uniq, inv = [], []
  for i in ids_per_rank:
    u, v = np.unique(np.asarray(i, np.int64), return_inverse=True)   # sorted uniques (radix sort order)
    uniq.append(u)
    inv.append(v)
  rows_u, ids_in, meta = fused_forward(shards, bucket_size, uniq)
  wire_ids = sum(len(u) for u in uniq)
  return [rows_u[q][inv[q]] for q in range(W)], wire_ids


def fused_backward_owner_order(ids_in, meta):
  """The order in which an owner sums row gradients: received position order
  (source rank major, then the requester's partitioned order)."""
  return [np.arange(len(a)) for a in ids_in]
