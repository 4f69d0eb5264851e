"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections
import csv
import re
import sys


def main(path, steps=3):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        t = t / 1e6 if unit == "ns" else t / 1e3 if unit == "us" else t
        key = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print(f"total {tot:.3f} ms over {steps} steps = {tot / steps:.3f} ms/step (serialised, cold-cache: compare shares)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / steps:9.3f} ms/step {n / steps:6.1f}x/step {100 * t / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
