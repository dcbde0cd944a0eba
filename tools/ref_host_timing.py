"""Reference kernels under (a) the reference's compiled host code and (b) its restatement, from the SAME start state
(the reference's own bootstrap), plus the restatement from bench.py's synthetic bootstrap: where does the time of the
reference arm go?  usage: python tools/ref_host_timing.py  (GPU box)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import ffi  # noqa: E402
import oracle_host  # noqa: E402
import torch  # noqa: E402

win, boot = bench.make_inputs(0)
args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])
R0 = np.asarray(boot[0], np.float32).reshape(3, 3)
epi = (R0, (R0.T.astype(np.float64) @ np.asarray(boot[1], np.float64)).astype(np.float32))
rboot = oracle_host.reference_host_bootstrap("ref", *args, epipolar=epi, config=bench.CONFIG)
print("closed-form depth / synthetic bootstrap depth: median %.4f, |log ratio| 90th pct %.3f" %
      (np.median(rboot[2] / boot[2]), np.percentile(np.abs(np.log(rboot[2] / boot[2])), 90)))
for rep in range(2):
    for name, fn in (
        ("compiled reference host, its own bootstrap", lambda: oracle_host.run_reference_host("ref", *args, config=bench.CONFIG, epipolar=epi)),
        ("restated host, the reference's bootstrap state", lambda: oracle_host.run_window("ref", *args, config=bench.CONFIG, boot=rboot)),
        ("restated host, bench.py's synthetic bootstrap", lambda: oracle_host.run_window("ref", *args, config=bench.CONFIG, boot=boot)),
    ):
        ffi.libc_srand(1000 + rep)
        torch.cuda.synchronize()
        t0 = time.time()
        r = fn()
        torch.cuda.synchronize()
        print(f"{name}: {time.time() - t0:.2f} s, registered {r['n_registered']}, stats {r.get('stats_ms')}")
