"""Group the SASS lines of one `ncu --page source --csv` export by their execution count: lines that execute equally often
belong to the same loop nest, which is enough to tell the per-candidate body of a kernel from its inner loops.
usage: python tools/summarize_source_regions.py gpurun_out/r02o/localprop_source.csv > profiles/r02_localprop_regions.md"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    kernel = rows[0][1]
    hdr, data = rows[1], rows[2:]
    ci, si = hdr.index("Instructions Executed"), hdr.index("Source")
    total = sum(int(r[ci]) for r in data)
    groups = collections.defaultdict(list)
    for r in data:
        groups[int(r[ci])].append(r[si].strip())
    print(f"kernel: `{kernel}`\n")
    print(f"{len(data)} SASS instructions, {total:,} executed warp instructions in this launch\n")
    print("| executions per SASS line | SASS lines | executed warp instructions | share | opcode mix of the region (static) |")
    print("|---:|---:|---:|---:|---|")
    for cnt, lines in sorted(groups.items(), key=lambda kv: -kv[0] * len(kv[1])):
        if cnt * len(lines) < 0.002 * total:
            continue
        mix = collections.Counter()
        for l in lines:
            tok = l.split()
            op = tok[1] if tok[0].startswith("@") else tok[0]
            mix[op.split(".")[0]] += 1
        top = ", ".join(f"{o} {n}" for o, n in mix.most_common(10))
        print(f"| {cnt:,} | {len(lines)} | {cnt * len(lines):,} | {cnt * len(lines) / total:.3f} | {top} |")


if __name__ == "__main__":
    main()
