#!/usr/bin/env python
"""Per-source-line stall samples and executed instructions from a .ncu-rep captured with --import-source on.

usage: python profiles/ncu_lines.py report.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    func, agg = None, {}
    for r in csv.reader(out.splitlines()):
        if r and r[0] == 'Function Name':
            func = r[1][:70]
            agg.setdefault(func, [])
        elif len(r) > 8 and r[0].strip().isdigit() and r[2] == '-':
            try:
                agg[func].append((int(r[0]), int(r[6]), int(r[7]), r[1]))
            except ValueError:
                pass
    for f, rows in agg.items():
        ts, ti = sum(x[1] for x in rows) or 1, sum(x[2] for x in rows) or 1
        print('=====', f, 'samples', ts, 'warp instructions', ti)
        for ln, s, i, src in sorted(rows, key=lambda x: -x[1])[:top]:
            print('%5d %6.2f%% smp %6.2f%% inst  %s' % (ln, 100.0 * s / ts, 100.0 * i / ti, src[:120]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
