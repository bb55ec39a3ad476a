#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path, top=45):
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    tot = 0.0
    for row in rows:
        name = row['Kernel Name'].split('(')[0][:56]
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        ms = v / 1e6 if unit in ('ns', 'nsecond') else v / 1e3 if unit in ('us', 'usecond') else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        tot += ms
    print('%-58s %5s %10s %9s %6s' % ('kernel', 'n', 'total ms', 'ms/launch', 'share'))
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-58s %5d %10.3f %9.4f %5.1f%%' % (k, n, ms, ms / n, 100 * ms / tot))
    print('total %.3f ms over %d launches' % (tot, len(rows)))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
