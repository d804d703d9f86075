"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python scripts/launch_agg.py launches.csv [divide_by]"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
    div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    h = rows[0]
    c = {k: i for i, k in enumerate(h)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[c['Metric Value']].replace(',', ''))
        except ValueError:
            continue
        u = r[c['Metric Unit']]
        v = v / 1e3 if u.startswith('n') else (v * 1e3 if u.startswith('m') else v)
        a = agg.setdefault(r[c['Kernel Name']][:100], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"total {tot / div:.1f} us per step ({len(rows) - 1} launches / {div:g})")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[1] / tot * 100:5.1f}%  n/step={a[0] / div:5.1f}  avg={a[1] / a[0]:8.1f}us  per-step={a[1] / div:8.1f}us  {k}")


if __name__ == "__main__":
    main()
