"""Group an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list by kernel: count, total, mean, share.
usage: python tools/launch_summary.py gpurun_out/launches.csv [divide_by_iterations]"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, iters=1):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rd:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn])
        name = re.sub(r"^void ", "", name)
        v = float(r[mv].replace(",", ""))
        u = r[mu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print(f"{n} launches, {total:.1f} us total" + (f" over {iters} iterations: {n / iters:.0f} launches, {total / iters:.1f} us per iteration" if iters > 1 else ""))
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / iters:10.1f} us {c / iters:6.0f}x {t / c:9.1f} us/launch {100 * t / total:5.1f}%  {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
