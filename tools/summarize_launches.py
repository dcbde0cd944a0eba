"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            name, v, unit = row["Kernel Name"], float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        except (KeyError, ValueError):
            continue
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
        name = re.sub(r"(vb::)?<?unnamed>::", "", name)
        a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = min(a[2], v)
        a[3] = max(a[3], v)
    return agg


def main():
    agg = load(sys.argv[1])
    tot = sum(a[1] for a in agg.values())
    print(f"source: `{sys.argv[1]}` — {sum(a[0] for a in agg.values())} launches, {tot / 1000:.2f} ms of kernel time "
          "(ncu per-launch durations: cold cache, serialised — compare shares, not absolutes)\n")
    print("| kernel | launches | sum µs | avg µs | min µs | max µs | share |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[2]:.1f} | {a[3]:.1f} | {a[1] / tot:.3f} |")


if __name__ == "__main__":
    main()
