"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [top_n]"""
import collections
import csv
import re
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (ValueError, KeyError):
            continue
        name = row['Kernel Name'].replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
        name = re.sub(r'\(.*', '', name)
        name = re.sub(r'<.*', '', name).replace('void ', '')
        unit = row['Metric Unit']
        v = v / 1e6 if unit == 'ns' else v / 1e3 if unit == 'us' else v
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f'# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.2f} ms total (cold-cache, serialised: compare shares)')
    print(f'{"ms":>10} {"n":>6} {"share":>7}  kernel')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{t:10.3f} {n:6d} {100 * t / tot:6.1f}%  {k[:100]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
