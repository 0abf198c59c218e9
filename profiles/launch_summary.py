"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'][:64]
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1.0)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:10.1f} us {v[0]:5d}x {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0] / 1e3:8.1f} us  {k}")


if __name__ == '__main__':
    main(sys.argv[1])
