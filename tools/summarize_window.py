"""Per-kernel table (launches, time, executed warp instructions) of one window captured with
`ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --csv`, plus the issue roofline of the whole window.
usage: python tools/summarize_window.py gpurun_out/rNN/window_launches.csv > profiles/rNN_window_launches.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(rows):
        try:
            name, m, v, u = r["Kernel Name"], r["Metric Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        except (KeyError, ValueError):
            continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
        name = re.sub(r"(vb::)?<?unnamed>::", "", name)
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "inst": 0.0})
        if m.startswith("gpu__time"):
            a["n"] += 1
            a["us"] += v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        else:
            a["inst"] += v
    tot_us = sum(a["us"] for a in agg.values())
    tot_i = sum(a["inst"] for a in agg.values())
    out = ["# One 30-iteration C2 window (640x480x8, 8192 hypotheses): every launch, duration and executed warp instructions", "",
           f"source: `ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none python tools/profile_window.py --iters 30` ({path}; profiler replay, serialised, cold caches: shares, not bench values)", "",
           "| kernel | launches | sum ms | avg µs | time share | warp instructions | instruction share |", "|---|---:|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out.append(f"| `{k}` | {a['n']} | {a['us'] / 1000:.2f} | {a['us'] / max(a['n'], 1):.1f} | {a['us'] / tot_us:.3f} | {a['inst']:.3e} | {a['inst'] / tot_i:.3f} |")
    out.append(f"| **total** | {sum(a['n'] for a in agg.values())} | {tot_us / 1000:.2f} | | 1 | {tot_i:.4e} | 1 |")
    peak = 148 * 4 * 1.965e9
    out += ["", f"Issue roofline of the whole window: {tot_i:.3e} warp instructions / (148 SMs x 4 schedulers x 1.965 GHz = {peak:.3e} /s) = "
            f"{tot_i / peak * 1e3:.1f} ms per window if every issue slot were used = {30 / (tot_i / peak):.0f} EM-iterations/s."]
    print("\n".join(out))


if __name__ == "__main__":
    main()
