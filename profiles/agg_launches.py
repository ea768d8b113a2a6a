"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (profiles helper)."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name, v, unit = row['Kernel Name'], float(row['Metric Value'].replace(',', '')), row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        m = re.search(r'(pwa::\w+|\w*layer_norm\w*|GammaBeta\w*|nvjet\w*|cutlass\w*|gemm\w*|\w+_kernel\w*)', name)
        short = (m.group(1) if m else name)[:60]
        a = agg.setdefault(short, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
        tot += v
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
    for k, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  max={mx:8.1f}  {k}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
