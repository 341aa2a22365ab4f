"""Generate the committed golden fixtures under tests/golden/.

Run HERE (the container with /root/reference): the partition / dual-modulo /
murmur3 vectors are produced by the reference's OWN CPU functors compiled in
place by oracle/build_ref.sh (oracle/_ref/libhbref.so); the alltoallv vectors
are transcribed from the reference's test file.  /root/reference does not exist
on the GPU box, so the tests read only the .npz/.json written here.

  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import hb_oracle as o  # noqa: E402


def main():
  assert o.ref() is not None, 'oracle/_ref not built (needs /root/reference)'
  out = {}
  # inputs follow hybridbackend/tensorflow/distribute/tests/partition_test.py:41-47,
  # :84-93 (np.random.seed(0), randint(-1e9, 1e9)), sizes reduced to keep the
  # fixture small; plus unsigned and edge cases.
  np.random.seed(0)
  cases = []
  x32 = np.random.randint(low=-1000000000, high=1000000000, size=10000, dtype=np.int32)
  cases.append(('i32_p5', x32, 5))
  np.random.seed(0)
  x64 = np.random.randint(low=-1000000000, high=1000000000, size=5000, dtype=np.int64)
  cases.append(('i64_p3', x64, 3))
  cases.append(('i64_p8', x64, 8))
  cases.append(('i64_p1', x64[:100], 1))
  cases.append(('u32_p7', (x32.astype(np.int64) + 3000000000).astype(np.uint32), 7))
  cases.append(('u64_p6', (x64.astype(np.uint64) * np.uint64(7919) + np.uint64(2**63)), 6))
  cases.append(('i64_empty_p7', np.array([], np.int64), 7))
  cases.append(('i64_big', np.array([2**63 - 1, -2**63, 0, -1, 1, 2**40 + 3, -(2**40) - 3] * 9,
                                    np.int64), 5))
  for name, x, p in cases:
    y, s, i = o.ref_partition_by_modulo(x, p)
    out[f'mod/{name}/x'] = x
    out[f'mod/{name}/p'] = np.array(p)
    out[f'mod/{name}/y'] = y
    out[f'mod/{name}/sizes'] = s
    out[f'mod/{name}/idx'] = i
  for name, x, p, m in [('i64_p4_m2', x64, 4, 2), ('i32_p2_m4', x32[:3000], 2, 4),
                        ('u64_p3_m3', out['mod/u64_p6/x'][:2000], 3, 3)]:
    for stage in (1, 2):
      y, s, i = o.ref_partition_by_dual_modulo(x, p, m, stage)
      key = f'dual/{name}/s{stage}'
      out[f'{key}/x'] = x
      out[f'{key}/pm'] = np.array([p, m, stage])
      out[f'{key}/y'] = y
      out[f'{key}/sizes'] = s
      out[f'{key}/idx'] = i
  keys = np.array([0, 1, -1, 42, 2**31, 2**32 + 7, -2**63, 2**63 - 1, 123456789012345], np.int64)
  out['murmur/keys'] = keys
  out['murmur/hash'] = np.array([o.ref_murmur3_hash32(int(k)) for k in keys], np.uint32)
  np.savez_compressed(os.path.join(HERE, 'partition_ref.npz'), **out)

  # hybridbackend/tensorflow/distribute/tests/alltoall_test.py (line numbers cited)
  golden = {
      'alltoallv': {  # :219-226
          'ids': [[1, 2, 3], [4, 5, 6]], 'sizes': [[1, 2], [1, 2]],
          'out_ids': [[1, 4], [2, 3, 5, 6]], 'out_sizes': [[1, 1], [2, 2]]},
      'alltoallv_n': {  # :254-269 (the only enabled case)
          'inputs': {'0': [{'ids': [1., 2., 3.], 'sizes': [1, 2]}, {'ids': [4., 5., 6.], 'sizes': [2, 1]}],
                     '1': [{'ids': [7., 8., 9.], 'sizes': [2, 1]}, {'ids': [10., 11., 12.], 'sizes': [1, 2]}]},
          'outputs': {'0': [{'ids': [1., 7., 8.], 'sizes': [1, 2]}, {'ids': [4., 5., 10.], 'sizes': [2, 1]}],
                      '1': [{'ids': [2., 3., 9.], 'sizes': [2, 1]}, {'ids': [6., 11., 12.], 'sizes': [1, 2]}]}},
      'alltoallv_grad': {  # :228-243: loss = mean(outputs), upstream g
          'g': 2.0, 'sizes': [[5, 1], [3, 4]]},
      'alltoallv_n_grad': {  # :288-304
          'g': 2.0, 'expected': 0.666667},
      'alltoall_grad': {  # :207-217: w=2,h=10,g=2 -> g/(w*h)
          'w': 2, 'h': 10, 'g': 2.0},
  }
  with open(os.path.join(HERE, 'alltoall_ref.json'), 'w') as f:
    json.dump(golden, f, indent=1)
  print('wrote', os.listdir(HERE))


if __name__ == '__main__':
  main()
