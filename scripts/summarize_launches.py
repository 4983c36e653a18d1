"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms (ncu-serialised, cold cache)")
    print(f"# {'ms':>10} {'launches':>8} {'share':>6} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.3f} {v[0]:8d} {100 * v[1] / tot:5.1f}% {1e3 * v[1] / v[0]:9.1f}  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
