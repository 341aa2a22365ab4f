#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an
.ncu-rep captured with --import-source on (built with -lineinfo).

  python profiles/srcprof.py gpurun_out/prof.ncu-rep <kernel regex> [top N] [launch idx]
"""
import csv
import io
import subprocess
import sys


def main():
  rep, kern = sys.argv[1], sys.argv[2]
  top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
  skip = sys.argv[4] if len(sys.argv) > 4 else '0'
  out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass',
                        '--kernel-name', f'regex:{kern}', '--launch-skip', skip, '--launch-count', '1'],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  fname, hdr, lines = None, None, []
  for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
      fname = r[1].split('/')[-1]
    elif len(r) > 5 and r[0] == 'Line No':
      hdr = r
    elif hdr is not None and len(r) == len(hdr) and r[0] not in ('', 'Line No'):
      d = dict(zip(hdr, r))
      try:
        lines.append((fname, int(r[0]), r[1].strip()[:90], int(d['Instructions Executed']),
                      int(d['Warp Stall Sampling (All Samples)'])))
      except ValueError:
        pass
  ti = sum(l[3] for l in lines) or 1
  ts = sum(l[4] for l in lines) or 1
  print(f'total warp-instructions {ti}, stall samples {ts}')
  print('--- top by instructions executed')
  for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f'{100*l[3]/ti:5.1f}% inst {100*l[4]/ts:5.1f}% stall  {l[0]}:{l[1]}  {l[2]}')
  print('--- top by stall samples')
  for l in sorted(lines, key=lambda x: -x[4])[:top]:
    print(f'{100*l[4]/ts:5.1f}% stall {100*l[3]/ti:5.1f}% inst  {l[0]}:{l[1]}  {l[2]}')


if __name__ == '__main__':
  main()
