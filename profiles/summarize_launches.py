"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v * 1e6 if u == "s" else v
        agg[row["Kernel Name"].split("(")[0]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'mean us':>10s} {'min us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:60]:60s} {len(v):8d} {sum(v) / len(v):10.2f} {min(v):10.2f} {100 * sum(v) / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
