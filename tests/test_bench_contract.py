"""bench.py contract checks that need no GPU: the reference arm prints one valid
JSON line, and the committed round-1 bench line carries every key the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
             'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config']


def test_reference_arm_prints_one_json_line():
  out = subprocess.check_output(
      [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
       '--warmup', '1', '--max-rows', '20000', '--batch', '2048'], text=True, timeout=300)
  lines = [l for l in out.strip().splitlines() if l.startswith('{')]
  assert len(lines) == 1
  d = json.loads(lines[0])
  for k in BASE_KEYS:
    assert k in d, k
  assert d['impl'] == 'reference' and d['value'] > 0 and d['higher_is_better'] is True
  assert d['metric'] == 'pooled-embedding-rows/sec' and 'workload' in d['config']
  cb = d['cpu_baseline']
  assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
  assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
  assert d['e2e']['value'] == d['value']


def test_reference_arm_other_ranks_exit_quietly():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
  out = subprocess.check_output(
      [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
      text=True, timeout=120, env=env)
  assert out.strip() == ''


import pytest


@pytest.mark.parametrize('name', ['bench_r1_n1.json', 'bench_r2_n1.json'])
def test_committed_bench_line_has_contract_keys(name):
  d = json.load(open(os.path.join(ROOT, 'profiles', name)))
  for k in BASE_KEYS + ['clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline']:
    assert k in d, k
  r = d['roofline']
  for k in ['bound', 'achieved', 'peak', 'unit', 'frac', 'traffic']:
    assert k in r, k
  assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
  assert d['gpu_launches'] > 0 and d['n_gpus'] == 1 and d['dtype'] == 'f32'
  e = d['e2e']
  assert e['h2d_bytes_per_step'] == 26 * 65536 * 8 and e['d2h_bytes_per_step'] == 65536 * 26 * 32 * 4
  assert e['value'] < d['value']            # e2e includes the PCIe copies
  cb = d['cpu_baseline']
  assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['sample']
  assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}


def test_round2_line_carries_the_second_data_points():
  d = json.load(open(os.path.join(ROOT, 'profiles', 'bench_r2_n1.json')))
  u = d['extra']['uniform']
  assert u['ms_per_step'] > 0 and 0 < u['lookup_fwd_roofline']['frac'] < 1.5
  v = d['e2e_variants']['device_out']
  assert v['d2h_bytes_per_step'] == 4 and v['h2d_bytes_per_step'] == 26 * 65536 * 8
  assert d['e2e']['value'] < v['value'] < d['value']
  names = {r['kernel'] for r in d['roofline_all']}
  assert {'lookup_fwd', 'sparse_update', 'sparse_update_long', 'sort_pass'} <= names
  n8 = json.load(open(os.path.join(ROOT, 'profiles', 'bench_r2_n8.json')))
  assert n8['n_gpus'] == 8 and n8['extra']['owner_load_unique_rows']['max_over_mean'] < 1.1
  assert n8['extra']['dedup']['wire_reduction'] > 3
