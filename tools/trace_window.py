"""Timeline of one C2 window (CUPTI through torch.profiler, nsys is not in this image): GPU busy time, idle gaps and
which kernel precedes them.  usage: python tools/trace_window.py [--iters 30] [--out gpurun_out/trace_r01.json]"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import synth  # noqa: E402
import voldor_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace.json"))
a = ap.parse_args()
win = synth.make_window(640, 480, 8, seed=100)
boot = (win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05, seed=7))
voldor_b200.set_bootstrap_override(*boot)
cfg = f"--silent --max_iters {a.iters} --no_trunc_iters 1000 --n_poses_to_sample 8192"
args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])
for _ in range(2):
    voldor_b200.voldor_ex(*args, config=cfg)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    r = voldor_b200.voldor_ex(*args, config=cfg)
    torch.cuda.synchronize()
prof.export_chrome_trace(a.out)
ev = [e for e in json.load(open(a.out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
busy = sum(e["dur"] for e in ev)
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
gap_after = collections.defaultdict(lambda: [0, 0.0])
per_kernel = collections.defaultdict(lambda: [0, 0.0])
for i, e in enumerate(ev):
    name = e["name"].replace("vb::(anonymous namespace)::", "").replace("void ", "").split("(")[0]
    per_kernel[name][0] += 1
    per_kernel[name][1] += e["dur"]
    if i + 1 < len(ev):
        g = ev[i + 1]["ts"] - (e["ts"] + e["dur"])
        if g > 0:
            gap_after[name][0] += 1
            gap_after[name][1] += g
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
union, (cs, ce) = 0.0, iv[0]
for a0, b0 in iv[1:]:
    if a0 > ce:
        union += ce - cs
        cs, ce = a0, b0
    else:
        ce = max(ce, b0)
union += ce - cs
print(f"window wall {r['stats_ms'][0]:.2f} ms; GPU activity span {span / 1000:.2f} ms, sum of activities {busy / 1000:.2f} ms, "
      f"GPU busy (union) {union / 1000:.2f} ms, idle {(span - union) / 1000:.2f} ms; {len(ev)} activities")
print("busy by kernel (us):")
for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {k[:60]:60s} n={v[0]:5d} sum={v[1]:9.1f} avg={v[1] / v[0]:7.1f}")
print("idle after (us):")
for k, v in sorted(gap_after.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {k[:60]:60s} n={v[0]:5d} sum={v[1]:9.1f} avg={v[1] / v[0]:7.1f}")
