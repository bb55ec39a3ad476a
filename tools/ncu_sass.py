#!/usr/bin/env python
"""Per-opcode executed-instruction counts and stall-sample totals of one kernel from an
`ncu --set full` report (SASS level; needs no source import).

    python tools/ncu_sass.py report.ncu-rep <kernel regex> [launch index]"""
import collections
import csv
import subprocess
import sys


def main(rep, pattern, which=0):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pattern],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'rows': [], 'hdr': None}
            blocks.append(cur)
        elif cur is not None and row and row[0] == 'Address':
            cur['hdr'] = row
        elif cur is not None and cur['hdr'] and len(row) == len(cur['hdr']):
            cur['rows'].append(row)
    b = blocks[which]
    h = b['hdr']
    iex, isrc, ismp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
    ops = collections.Counter()
    smp = collections.Counter()
    stalls = collections.Counter()
    tot = 0
    for r in b['rows']:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
        op = op.split('.')[0] + ('.' + '.'.join(op.split('.')[1:3]) if op.startswith(('F2F', 'I2F', 'F2I', 'LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED')) else '')
        n = int(r[iex])
        ops[op] += n
        smp[op] += int(r[ismp])
        tot += n
        for i, nme in stall_cols:
            stalls[nme] += int(r[i])
    print(b['name'][:100])
    print('launches in report matching: %d; warp instructions executed: %d; SASS lines %d' % (len(blocks), tot, len(b['rows'])))
    print('%-22s %12s %6s %8s' % ('opcode', 'executed', 'share', 'samples'))
    for op, n in ops.most_common(28):
        print('%-22s %12d %5.1f%% %8d' % (op, n, 100.0 * n / tot, smp[op]))
    ts = sum(stalls.values())
    print('stall samples:', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / ts) for k, v in stalls.most_common(8)))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
