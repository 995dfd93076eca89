#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel of an
.ncu-rep captured with --import-source on:
  python profiles/source_hotspots.py rep.ncu-rep <kernel regex> [top N]"""
import csv
import io
import subprocess
import sys


def main():
  rep, pat = sys.argv[1], sys.argv[2]
  top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
  out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name',
                        'regex:' + pat, '--print-source', 'cuda,sass'],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  fname, data, hdr, seen_fn = None, [], None, 0
  for r in rows:
    if not r:
      continue
    if r[0] == 'File Path':
      fname = r[1].split('/')[-1]
    elif r[0] == 'Function Name':
      seen_fn += 1
    elif r[0] == 'Line No':
      hdr = r
    elif hdr is not None and r[0].isdigit():
      iex = hdr.index('Instructions Executed')
      ist = hdr.index('Warp Stall Sampling (All Samples)')
      num = lambda v: int(v) if v.isdigit() else 0
      data.append((fname, int(r[0]), r[1].strip(), num(r[iex]), num(r[ist])))
  tot = sum(d[3] for d in data) or 1
  st = sum(d[4] for d in data) or 1
  print('kernel regex %s: %d instructions executed, %d stall samples (all captured launches)' % (pat, tot, st))
  for f, ln, src, e, t in sorted(data, key=lambda d: -d[3])[:top]:
    print('%-14s %4d  inst %5.2f%%  stall %5.2f%%  %s' % (f, ln, 100.0 * e / tot, 100.0 * t / st, src[:96]))


if __name__ == '__main__':
  main()
