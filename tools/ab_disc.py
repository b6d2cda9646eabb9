#!/usr/bin/env python
"""Developer A/B harness: run the default bench once per configuration, round-robin, and print the median stage times.
Usage: ab_disc.py [--rounds R] "NAME=VALUE,NAME2=VALUE2" "NAME=OTHER" ...   (an empty string = the default build)."""
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
rounds = 3
if args and args[0] == "--rounds":
    rounds = int(args[1])
    args = args[2:]
configs = args or [""]
res = {c: [] for c in configs}
for r in range(rounds):
    for c in configs:
        env = dict(os.environ)
        for kv in filter(None, c.split(",")):
            k, v = kv.split("=", 1)
            env[k] = v
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "30", "--warmup", "5", "--no-cpu-baseline"],
                             capture_output=True, text=True, env=env, timeout=300)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if not line:
            print(repr(c), "FAILED", out.stderr[-300:], flush=True)
            continue
        d = json.loads(line[-1])
        res[c].append((d["stage_ms"]["disc_reward"], d["stage_ms"]["sample+obs"], d["ms_per_step"], d["e2e"]["value"]))
        print(f"round {r} {c!r}: disc {res[c][-1][0]:.4f} ms  obs {res[c][-1][1]:.4f} ms  step {res[c][-1][2]:.4f} ms", flush=True)
for c, v in res.items():
    if v:
        print(f"MEDIAN {c!r}: disc {statistics.median(x[0] for x in v):.4f}  obs {statistics.median(x[1] for x in v):.4f}  "
              f"step {statistics.median(x[2] for x in v):.4f}  e2e {statistics.median(x[3] for x in v):.4g}")
