"""Summarise an ncu launch-list CSV (gpu__time_duration.sum) per kernel: count, total, mean, share."""
import collections
import csv
import re
import sys


def main(path, skip_frac=0.5):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    data = rows[hi + 1:]
    data = data[int(len(data) * skip_frac):]
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        full = r[kn]
        name = re.sub(r"\(.*", "", full)
        t = float(r[mv].replace(",", "")) / 1000
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print(f"{len(data)} launches, {tot:.1f} us total")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.1f} us {v[0]:4d}x {v[1] / v[0]:7.1f} us/launch {100 * v[1] / tot:5.1f}%  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
