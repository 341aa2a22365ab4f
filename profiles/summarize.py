"""Turn the ncu outputs a gpurun call brought back (gpurun_out/) into the small,
committed summaries under profiles/:
  <tag>_launches.md   every hb:: kernel of the launch list with count / avg / share
  <tag>_ncu.md        key `--set full` metrics of the captured hb:: kernels
  traffic.json        dram read+write bytes per launch, read by bench.py

  python profiles/summarize.py r1 gpurun_out/launches_r1.csv gpurun_out/prof_r1.ncu-rep
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEY = {'hb::lookup_fwd_kernel': 'lookup_fwd', 'hb::sparse_update_kernel': 'sparse_update',
       'hb::sparse_update_fixup_kernel': 'sparse_update_fixup', 'hb::bucket_pass_kernel': 'sort_pass',
       'hb::bucket_hist_kernel': 'sort_hist', 'hb::sh_owner_gather_kernel': 'sharded_owner_gather',
       'hb::sh_push_grads_kernel': 'sharded_push_grads',
       # round 2
       'hb::update_short_kernel': 'sparse_update', 'hb::update_long_kernel': 'sparse_update_long',
       'hb::cluster_sort_runs_kernel': 'sort_pass', 'update_short_kernel': 'sparse_update',
       'update_long_kernel': 'sparse_update_long', 'cluster_sort_runs_kernel': 'sort_pass',
       'lookup_fwd_kernel': 'lookup_fwd'}


def launches(tag, path):
  lines = [l for l in open(path) if not l.startswith('==')]
  agg = collections.defaultdict(list)
  for row in csv.DictReader(lines):
    try:
      v = float(row['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
      continue
    u = row.get('Metric Unit', 'ns')
    v = v / 1000 if u in ('ns', 'nsecond') else (v * 1000 if u in ('ms', 'msecond') else v)
    agg[row['Kernel Name'].split('(')[0]].append(v)
  ours = {k: v for k, v in agg.items() if 'hb::' in k}
  tot = sum(sum(v) for v in ours.values())
  out = [f'# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)', '',
         'Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.',
         'Only hb:: kernels are listed (torch setup kernels of bench.py excluded).', '',
         '| kernel | launches | avg us | total us | share of hb:: time |', '|---|---|---|---|---|']
  for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
    out.append(f'| `{k.replace("void ", "")}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {sum(v) / tot * 100:.1f}% |')
  open(os.path.join(HERE, f'{tag}_launches.md'), 'w').write('\n'.join(out) + '\n')


def full(tag, rep):
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units, data = rows[0], rows[1], rows[2:]
  want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'launch__grid_size', 'lts__t_sector_hit_rate.pct',
          'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
          'smsp__warp_issue_stalled_barrier_per_warp_active.pct']
  idx = [(w, hdr.index(w)) for w in want if w in hdr]
  ki = hdr.index('Kernel Name')
  out = [f'# {tag}: ncu --set full --clock-control none (one row per captured launch)', '',
         '| kernel | ' + ' | '.join(w for w, _ in idx) + ' |', '|---|' + '---|' * len(idx)]
  traffic = collections.defaultdict(list)

  def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
  for d in data:
    name = d[ki]
    out.append(f'| `{name[:70]}` | ' + ' | '.join(f'{d[i]} {units[i]}' for _, i in idx) + ' |')
    for k, short in KEY.items():
      if k.replace('hb::', '') in name:
        ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        traffic[short].append(to_bytes(d[ir], units[ir]) + to_bytes(d[iw], units[iw]))
  open(os.path.join(HERE, f'{tag}_ncu.md'), 'w').write('\n'.join(out) + '\n')
  tj = {k: max(v) for k, v in traffic.items()}  # the largest launch of a kernel family
  tj['_source'] = f'{tag}: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)'
  json.dump(tj, open(os.path.join(HERE, 'traffic.json'), 'w'), indent=1)


if __name__ == '__main__':
  tag = sys.argv[1]
  launches(tag, sys.argv[2])
  if len(sys.argv) > 3:
    full(tag, sys.argv[3])
  print(os.listdir(HERE))
