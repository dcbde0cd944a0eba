"""Condense one `ncu --set full` report into the handful of numbers the design discussion uses.
usage: python tools/summarize_ncu.py gpurun_out/search_r01.ncu-rep profiles/r01_ncu_search  (writes .md and .json)
       the first argument may also be the `ncu -i x.ncu-rep --page raw --csv` export of a report"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1/TEX hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("sass__inst_executed_local_loads", "local-memory loads"),
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in data:
        d = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("unnamed>::", "").replace("void ", "")}
        for key, _ in WANT:
            if key in col:
                d[key] = (float(r[col[key]].replace(",", "")), units[col[key]])
        launches.append(d)
    md = [f"source: `{rep}` (ncu --set full --clock-control none; profiler replay — durations are not bench values)\n"]
    md.append("| metric | " + " | ".join(f"launch {i}" for i in range(len(launches))) + " |")
    md.append("|---|" + "---:|" * len(launches))
    md.append("| kernel | " + " | ".join(f"`{l['kernel']}`" for l in launches) + " |")
    for key, label in WANT:
        if key in launches[0]:
            md.append(f"| {label} (`{key}`) | " + " | ".join(f"{l[key][0]:,.2f} {l[key][1]}" for l in launches) + " |")
    js = {"source": rep, "kernel": launches[0]["kernel"], "launches": len(launches)}
    tr = []
    for l in launches:
        rd, wr = l["dram__bytes_read.sum"], l["dram__bytes_write.sum"]
        tr.append(rd[0] * UNIT_SCALE[rd[1]] + wr[0] * UNIT_SCALE[wr[1]])
    js["dram_bytes_per_launch"] = sum(tr) / len(tr)
    js["issue_active_pct"] = sum(l["smsp__issue_active.avg.pct_of_peak_sustained_active"][0] for l in launches) / len(launches)
    js["xu_pipe_pct"] = sum(l["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"][0] for l in launches) / len(launches)
    js["warp_instructions"] = sum(l["smsp__inst_executed.sum"][0] for l in launches) / len(launches)
    open(out + ".md", "w").write("\n".join(md) + "\n")
    json.dump(js, open(out + ".json", "w"), indent=1)
    print(json.dumps(js))


if __name__ == "__main__":
    main()
