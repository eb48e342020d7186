#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line:
executed warp instructions, stall samples and shared-memory wavefronts, per kernel launch.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_lines.py src.csv [top N]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    i = 0
    launch = -1
    while i < len(rows):
        r = rows[i]
        if r and r[0] == 'Function Name':
            fn = r[1]
            hdr = rows[i + 1]
            launch += 1
            col = {h: k for k, h in enumerate(hdr)}
            # the header has two "Source" columns: the first is the CUDA line, the second the SASS text
            src_cols = [k for k, h in enumerate(hdr) if h == 'Source']
            agg = defaultdict(lambda: [0, 0, 0, 0, ''])
            j = i + 2
            cur_file = rows[i - 1][1] if rows[i - 1] and rows[i - 1][0] == 'File Path' else ''
            tot = 0
            while j < len(rows) and rows[j] and rows[j][0] not in ('File Path', 'Function Name'):
                rr = rows[j]
                try:
                    ex = int(rr[col['Instructions Executed']] or 0)
                    smp = int(rr[col['# Samples']] or 0)
                    wf = int(rr[col['L1 Wavefronts Shared']] or 0)
                    wfi = int(rr[col['L1 Wavefronts Shared Ideal']] or 0)
                except (ValueError, KeyError):
                    j += 1
                    continue
                key = (rr[col['Line No']], rr[src_cols[0]].strip()[:90])
                a = agg[key]
                a[0] += ex
                a[1] += smp
                a[2] += wf
                a[3] += wfi
                tot += ex
                j += 1
            if tot:
                print('=== launch %d  %s  [%s]  total warp instructions %d' % (launch, fn[:60], cur_file[-30:], tot))
                for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
                    print('%8d %5.1f%%  smp %6d  wf %8d/%8d  L%-4s %s' % (a[0], 100.0 * a[0] / tot, a[1], a[2], a[3], ln, src))
            i = j
        else:
            i += 1


if __name__ == '__main__':
    main()
