"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share,
for the whole run and for the encoder region (launches before the first decoder kernel)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    order = []
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        unit = row['Metric Unit']
        v = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
        name = re.sub(r'^.*?::', '', re.sub(r'\(.*', '', row['Kernel Name']))
        order.append((name, v, row.get('Grid Size', '')))

    def table(rows, title):
        agg = collections.defaultdict(lambda: [0, 0.0])
        for n, v, _ in rows:
            agg[n][0] += 1
            agg[n][1] += v
        tot = sum(v[1] for v in agg.values())
        print(f'--- {title}: {len(rows)} launches, {tot / 1e3:.2f} ms (serialised, cold-cache)')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f'{k[:60]:60s} n={v[0]:6d} total={v[1] / 1e3:9.2f} ms avg={v[1] / v[0]:8.2f} us share={v[1] / tot * 100:5.1f}%')

    table(order, 'all')
    idx = [i for i, (n, _, _) in enumerate(order) if 'embed_ln' in n]
    if idx:
        table(order[:idx[0]], 'encoder region')
        table(order[idx[0]:], 'decoder region')


if __name__ == '__main__':
    main(sys.argv[1])
