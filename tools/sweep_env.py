#!/usr/bin/env python
"""Run bench.py under several values of an environment variable and print the per-kernel times."""
import json, os, subprocess, sys
var, vals = sys.argv[1], sys.argv[2].split(",")
extra = sys.argv[3:]
for v in vals:
    env = dict(os.environ, **{var: v})
    out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline"] + extra, env=env, capture_output=True, text=True).stdout
    try:
        d = json.loads(out.strip().splitlines()[-1])
        k = d["kernel_ms_per_step_rank0"]
        print(var, v, "minhash_ms=%.2f" % k["minhash_ms"], "dedup=%.2f ordered=%.2f filter=%.2f value=%.3f" % (k["hash_dedup_ms"], k["ordered_ms"], k["filter_ms"], d["value"]), flush=True)
    except Exception as e:
        print(var, v, "FAILED", e, out[-300:], flush=True)
