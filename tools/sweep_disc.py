#!/usr/bin/env python
"""Developer sweep: discriminator chunk size (tiles per CTA) x kernel variant (CTA pair / single) on the default bench."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for tiles in (2, 4, 8, 16, 32):
    for single in (0, 1):
        env = dict(os.environ, AMP_B200_DISC_TILES_PER_CTA=str(tiles), AMP_B200_DISC_PAIR=str(1 - single))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-cpu-baseline"],
                             capture_output=True, text=True, env=env, timeout=300)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if not line:
            print("tiles", tiles, "single", single, "FAILED", out.stderr[-300:])
            continue
        d = json.loads(line[-1])
        print("tiles", tiles, "single", single, "ms_per_step", round(d["ms_per_step"], 4), "disc_ms", round(d["stage_ms"]["disc_reward"], 4), flush=True)
