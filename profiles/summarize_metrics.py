"""Per-kernel evidence table from ONE `ncu --metrics <list> --csv` capture of a training step: for every kernel name the
launch count, total / average device time, DRAM bytes moved (read + write) and the achieved DRAM GB/s against the measured
HBM peak (MEASURED_PEAKS.json), tensor-pipe %, L2 hit rate, achieved warps. Times are cold-cache and serialised (compare
shares); DRAM bytes are exact.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,\
sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file X.csv <cmd>
  python profiles/summarize_metrics.py X.csv [top]
"""
import collections
import csv
import json
import os
import re
import sys

UNIT = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 's': 1e6, 'second': 1e6,
        'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def short(name):
    name = name.replace('void ', '').replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
    name = re.sub(r'\(.*$', '', name)
    return re.sub(r'\(int\)|\(bool\)', '', name)[:84]


def main(path, top=60):
    peak = 6576.4
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = float(json.load(open(pk))['hbm_gbs'])
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    per_launch = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row['ID'], short(row['Kernel Name']))
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        per_launch.setdefault(key, {})[row['Metric Name']] = v * UNIT.get(row['Metric Unit'], 1.0)
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for (_, name), m in per_launch.items():
        a = agg[name]
        t = m.get('gpu__time_duration.sum', 0.0)
        a['n'] += 1
        a['us'] += t
        a['bytes'] += m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)
        for src, dst in (('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor'),
                         ('lts__t_sector_hit_rate.pct', 'l2hit'),
                         ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps')):
            a[dst] += m.get(src, 0.0) * t           # time-weighted
    T = sum(a['us'] for a in agg.values())
    print(f'# {path}: {int(sum(a["n"] for a in agg.values()))} launches, {T / 1e3:.2f} ms serialised; HBM peak {peak:.0f} GB/s (measured)')
    print(f'{"us":>9} {"share":>6} {"n":>5} {"avg_us":>8} {"DRAM_MB":>9} {"GB/s":>7} {"%HBM":>5} {"tens%":>6} {"L2hit%":>6} {"warps%":>6}  kernel')
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['us'])[:top]:
        us = max(a['us'], 1e-9)
        gbs = a['bytes'] / us / 1e3
        print(f'{a["us"]:9.0f} {100 * a["us"] / T:5.1f}% {int(a["n"]):5d} {a["us"] / a["n"]:8.1f} {a["bytes"] / 1e6:9.1f} {gbs:7.0f} '
              f'{100 * gbs / peak:5.1f} {a["tensor"] / us:6.1f} {a["l2hit"] / us:6.1f} {a["warps"] / us:6.1f}  {name}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
