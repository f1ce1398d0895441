#!/usr/bin/env python3
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list per kernel function:
launches, total duration, share of the step, DRAM bytes.  tools/ncu_summary.py launches.csv [launches_per_step] > summary.txt
(launches_per_step: keep only the LAST that many launches = one whole forward when the run did several)."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    by_id = OrderedDict()
    for r in rows:
        d = by_id.setdefault(r['ID'], {'name': r['Kernel Name']})
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        m = r['Metric Name']
        if m == 'gpu__time_duration.sum':
            d['ns'] = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        elif m == 'dram__bytes_read.sum':
            d['rd'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        elif m == 'dram__bytes_write.sum':
            d['wr'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    launches = list(by_id.values())
    if last:
        launches = launches[-last:]
    agg = OrderedDict()
    for d in launches:
        name = re.sub(r'\(.*$', '', d['name']).replace('void ', '').replace('dlv3p::', '')
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get('ns', 0.0)
        a[2] += d.get('rd', 0.0)
        a[3] += d.get('wr', 0.0)
    total = sum(a[1] for a in agg.values())
    print('launches %d, sum of kernel durations %.3f ms, DRAM read %.1f MB, write %.1f MB' % (len(launches), total / 1e6, sum(a[2] for a in agg.values()) / 1e6,
                                                                                             sum(a[3] for a in agg.values()) / 1e6))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-60s launches %3d  %7.3f ms  %4.1f%%  dram_read_MB %9.1f  dram_write_MB %9.1f' % (name, a[0], a[1] / 1e6, 100 * a[1] / total, a[2] / 1e6, a[3] / 1e6))


if __name__ == '__main__':
    main()
