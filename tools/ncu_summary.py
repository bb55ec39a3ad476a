#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one line per launch with time, DRAM bytes, throughput
fractions, and a kernel -> DRAM bytes per launch table (JSON) for bench.py's roofline.traffic.

    python tools/ncu_summary.py report.ncu-rep out_summary.txt out_traffic.json "note"
"""
import csv
import json
import subprocess
import sys

COLS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'lts__t_sector_hit_rate.pct']
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
TSCALE = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3, 'second': 1e6, 's': 1e6}


def main(rep, out_txt, out_json, note=''):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in COLS]
    ik = hdr.index('Kernel Name')
    lines = ['ncu --set full --clock-control none, B200. ' + note, '',
             '%-32s %9s %10s %10s %7s %7s %7s %7s %7s %5s %7s %6s %6s' % (
                 'kernel', 'time us', 'dram rd MB', 'dram wr MB', 'dram%', 'sm%', 'issue%', 'fp64%', 'warps%', 'regs',
                 'grid', 'block', 'L2hit%')]
    traffic = {}
    for r in rows[2:]:
        name = r[ik].split('(')[0].replace('void ', '')
        rd = float(r[idx[1]].replace(',', '')) * SCALE[units[idx[1]]]
        wr = float(r[idx[2]].replace(',', '')) * SCALE[units[idx[2]]]
        t_us = float(r[idx[0]].replace(',', '')) * TSCALE[units[idx[0]]]
        v = [r[i] for i in idx[3:]]
        lines.append('%-32s %9.1f %10.1f %10.1f %7.1f %7.1f %7.1f %7.1f %7.1f %5s %7s %6s %6.1f' % (
            name[:32], t_us, rd / 1e6, wr / 1e6, float(v[0]), float(v[1]), float(v[2]), float(v[3]), float(v[4]),
            v[5], v[6], v[7], float(v[8])))
        traffic.setdefault(name.split('<')[0], rd + wr)
    with open(out_txt, 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    with open(out_json, 'w') as fh:
        json.dump(traffic, fh, indent=1)
    print('\n'.join(lines))


if __name__ == '__main__':
    main(*sys.argv[1:5])
