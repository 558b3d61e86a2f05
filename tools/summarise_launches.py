"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel name.

    python tools/summarise_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")
        name = re.sub(r"at::native::.*?(\w+_kernel\w*).*", r"torch::\1", name)
        agg[name][0] += 1
        agg[name][1] += float(r[14]) / 1e3
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised under ncu)")
    print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>9s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {n:8d} {us:12.1f} {100 * us / total:6.1f}% {us / n:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
