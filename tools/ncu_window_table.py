#!/usr/bin/env python
"""Per-kernel totals of a whole-window ncu launch list (tools/prof_window.sh): launches, device time, DRAM bytes."""
import csv, collections, sys, json
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
by = collections.OrderedDict()
per = {}
for row in csv.DictReader(lines):
    k = int(row['ID'])
    per.setdefault(k, {'name': row['Kernel Name']})[row['Metric Name']] = (float(row['Metric Value'].replace(',', '')), row['Metric Unit'])
def to_base(v, u):
    m = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1, 'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return v * m.get(u, 1)
names = {'0': 'FETCH', '1': 'HEAD', '2': 'FRONT', '3': 'SEARCH', '4': 'TAIL', '5': 'CORR'}
for k, d in per.items():
    n = d['name'].split('(')[0]
    if 'lgar_wave_kernel<' in n:
        n = 'lgar_wave_kernel<%s> (%s)' % (n.split('<')[1][0], names[n.split('<')[1][0]])
    a = by.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += to_base(*d['gpu__time_duration.sum'])
    a[2] += to_base(*d['dram__bytes_read.sum'])
    a[3] += to_base(*d['dram__bytes_write.sum'])
T = sum(a[1] for a in by.values()); R = sum(a[2] for a in by.values()); W = sum(a[3] for a in by.values())
print('| kernel | launches | total ms | share | DRAM read GB | DRAM write GB |\n|---|---:|---:|---:|---:|---:|')
for n, a in sorted(by.items(), key=lambda x: -x[1][1]):
    print('| `%s` | %d | %.2f | %.1f%% | %.2f | %.2f |' % (n, a[0], a[1] * 1e3, 100 * a[1] / T, a[2] / 1e9, a[3] / 1e9))
print('| total | %d | %.2f | 100%% | %.2f | %.2f |' % (sum(a[0] for a in by.values()), T * 1e3, R / 1e9, W / 1e9))
print(json.dumps({'window_ms_under_ncu': T * 1e3, 'dram_bytes_per_window': R + W, 'dram_bytes_per_bench_step': (R + W) / steps, 'steps': steps}))
