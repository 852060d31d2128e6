"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-launch times for one
timestep and per-kernel shares.  usage: python profiles/launch_table.py launches.csv [n_per_step]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            name = re.sub(r'^void |rvt::|\(.*', '', row['Kernel Name'])
            rows.append((name, row['Grid Size'].replace(' ', ''), float(row['Metric Value'].replace(',', '')) / 1e3))
    return rows


if __name__ == '__main__':
    rows = load(sys.argv[1])
    # the capture starts at process start: drop PyTorch's own kernels (weight packing, fills) and keep this library's
    rows = [r for r in rows if not (r[0].startswith('at::') or r[0].startswith('<unnamed>') or 'elementwise' in r[0])]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 52
    print(f'{len(rows)} launches; first {n}:')
    for name, grid, us in rows[:n]:
        print(f'  {name:42s} grid {grid:>14s} {us:9.1f} us')
    agg = collections.defaultdict(lambda: [0.0, 0])
    for name, _, us in rows:
        agg[name][0] += us
        agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f'total {tot:.1f} us over {len(rows)} launches')
    for name, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'  {us / tot * 100:5.1f}%  {us:9.1f} us  x{cnt:<4d} {name}')
