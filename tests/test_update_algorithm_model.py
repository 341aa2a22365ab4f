"""A scalar model of the tile / partial / chain bookkeeping of
hybridbackend_b200/csrc/sparse_update.cu (sparse_update_kernel + fix-up): warp tiles of
32 sorted entries, group tiles of G entries, super-tiles of 256, flags
FirstOpen / Both / LastOpen, in-CTA chains and cross-super-tile chains.  The model
mirrors the kernel's decisions one to one (same predicates), with integers for the
gradients, and must apply every distinct row exactly once with the full sum -- for
every G, ragged tails, runs ending exactly on tile borders and very long runs.
It is a design check that needs no GPU; the kernels themselves are checked by the
-m gpu parity tests."""
import numpy as np
import pytest

FIRST_OPEN, BOTH, LAST_OPEN = 1, 2, 4


def model_update(keys, vals, G, super_tile=256):
  n = len(keys)
  applied = {}          # key -> list of applied sums (must end with exactly one)

  def apply(k, s):
    applied.setdefault(int(k), []).append(int(s))

  def same_prev(e):     # global: entry e has the same key as e-1
    return e > 0 and keys[e] == keys[e - 1]

  groups = super_tile // G
  nst = (n + super_tile - 1) // super_tile
  st_flag = [0] * nst
  st_part = [[0, 0] for _ in range(nst)]
  st_key = [[None, None] for _ in range(nst)]
  for st in range(nst):
    flag = [0] * groups
    part = [[0, 0] for _ in range(groups)]
    pkey = [[None, None] for _ in range(groups)]
    for g in range(groups):
      t0 = st * super_tile + g * G
      cnt = max(0, min(G, n - t0))
      if cnt == 0:
        continue
      w0 = (t0 // 32) * 32
      wcnt = max(0, min(32, n - w0))
      warp_open_right = wcnt == 32 and w0 + 32 < n and keys[w0 + 32] == keys[w0 + 31]
      first_open_left = same_prev(t0)
      seen_tail = False
      acc = 0
      for j in range(cnt):
        e = t0 + j
        head = not same_prev(e)
        cont = (j != 0) and not head
        acc = acc + vals[e] if cont else vals[e]
        last = j == cnt - 1
        if e + 1 < w0 + wcnt:
          next_same = same_prev(e + 1)
        else:
          next_same = warp_open_right
        tail = (not next_same) or last
        if tail:
          ol = first_open_left and not seen_tail
          orr = last and next_same
          if not ol and not orr:
            apply(keys[e], acc)
          elif ol:
            part[g][0], pkey[g][0] = acc, keys[e]
            flag[g] |= FIRST_OPEN | (BOTH if orr else 0)
          else:
            part[g][1], pkey[g][1] = acc, keys[e]
            flag[g] |= LAST_OPEN
          seen_tail = True
    # in-CTA combine
    if flag[0] & FIRST_OPEN:
      acc, t, both = part[0][0], 0, True
      while True:
        if not flag[t] & BOTH:
          both = False
          break
        t += 1
        if t == groups or not flag[t] & FIRST_OPEN:
          break
        acc += part[t][0]
      st_part[st][0], st_key[st][0] = acc, pkey[0][0]
      st_flag[st] |= FIRST_OPEN | (BOTH if both else 0)
    for g in range(groups):
      if flag[g] & LAST_OPEN:
        acc, key, t, closed = part[g][1], pkey[g][1], g + 1, False
        while t < groups:
          if not flag[t] & FIRST_OPEN:
            break
          acc += part[t][0]
          if not flag[t] & BOTH:
            closed = True
            break
          t += 1
        if closed:
          apply(key, acc)
        else:
          st_part[st][1], st_key[st][1] = acc, key
          st_flag[st] |= LAST_OPEN
  # fix-up across super-tiles
  for st in range(nst):
    if st_flag[st] & LAST_OPEN:
      acc, key, t = st_part[st][1], st_key[st][1], st + 1
      while t < nst:
        if not st_flag[t] & FIRST_OPEN:
          break
        assert st_key[t][0] == key
        acc += st_part[t][0]
        if not st_flag[t] & BOTH:
          break
        t += 1
      apply(key, acc)
  return applied


def _check(keys, vals, G):
  applied = model_update(keys, vals, G)
  exp = {}
  for k, v in zip(keys, vals):
    exp[int(k)] = exp.get(int(k), 0) + int(v)
  assert set(applied) == set(exp)
  for k, sums in applied.items():
    assert len(sums) == 1, f'row {k} applied {len(sums)} times'
    assert sums[0] == exp[k]


@pytest.mark.parametrize('G', [1, 2, 4, 8, 16, 32])
def test_model_random(G):
  rng = np.random.RandomState(G)
  for trial in range(60):
    n = int(rng.choice([1, 2, 31, 32, 33, 255, 256, 257, 511, 777, 1024, 3000]))
    nkeys = int(rng.choice([1, 2, 3, 10, 100, 5000]))
    keys = np.sort(rng.randint(0, nkeys, n))
    vals = rng.randint(1, 1000, n)
    _check(keys, vals, G)


@pytest.mark.parametrize('G', [1, 4, 8, 32])
def test_model_runs_on_borders(G):
  # runs that end exactly at group / warp / super-tile borders, and one huge run
  for n, cuts in [(1024, [G, 32, 256, 512, 768]), (2048, [256]), (1000, [255, 257, 512, 999]),
                  (4096, []), (513, [512])]:
    keys = np.zeros(n, np.int64)
    for c in cuts:
      if c < n:
        keys[c:] += 1
    vals = np.arange(1, n + 1)
    _check(keys, vals, G)


def _fixup_round_model(flags, start):
  """sparse_update_fixup_kernel's ballot logic: how many super-tiles after `start`
  join the chain, walked 32 at a time."""
  nst = len(flags)
  t, total, done = start + 1, 0, False
  while not done and t < nst:
    fl = [(flags[t + i] if t + i < nst else 0) for i in range(32)]
    stop_a = [not (f & FIRST_OPEN) for f in fl]
    stop_b = [bool(f & FIRST_OPEN) and not (f & BOTH) for f in fl]
    pa = stop_a.index(True) if True in stop_a else 32
    pb = stop_b.index(True) if True in stop_b else 32
    if pb < pa:
      m, done = pb + 1, True
    else:
      m = pa
      done = pa < 32
    total += m
    t += 32
  return total


def _fixup_sequential(flags, start):
  t, total = start + 1, 0
  while t < len(flags):
    if not flags[t] & FIRST_OPEN:
      break
    total += 1
    if not flags[t] & BOTH:
      break
    t += 1
  return total


def test_fixup_round_logic_equals_sequential_walk():
  rng = np.random.RandomState(0)
  for length in [0, 1, 5, 31, 32, 33, 63, 64, 65, 200]:
    for tail in ['closed', 'data_end']:
      # chain head at 0, `length` continuing tiles (all BOTH except the closing one)
      flags = [LAST_OPEN] + [FIRST_OPEN | BOTH] * length
      if tail == 'closed':
        flags += [FIRST_OPEN] + [int(rng.choice([0, LAST_OPEN, FIRST_OPEN]))] * 3
      assert _fixup_round_model(flags, 0) == _fixup_sequential(flags, 0), (length, tail)
