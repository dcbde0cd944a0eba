"""bench.py output contract on the GPU-less leg (`--impl cpu-port`: the oracle's CPU port under the restated reference
orchestration — the fallback of the reference arm when oracle/_ref is absent)."""
import json
import os
import subprocess
import sys

import ffi


def test_reference_arm_line_has_every_contract_key():
    r = subprocess.run([sys.executable, os.path.join(ffi.ROOT, "bench.py"), "--impl", "cpu-port", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "EM-iters/sec" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["scaling"] == "weak" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == line["unit"] and cb["sample"]
