#!/usr/bin/env python
"""bench.py's main timed region + parity check under different svt_set_option switches, alternating, on one box.
    python tools/option_probe.py resid_bf16=0 resid_bf16=1 [rounds]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
opts = [a for a in sys.argv[1:] if "=" in a]
rounds = int(next((a for a in sys.argv[1:] if "=" not in a), "2"))
for r in range(rounds):
    for o in opts:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-aux",
                              "--no-cpu-baseline", "--no-e2e"], env=dict(os.environ, SVT_B200_OPTIONS=o), capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(f"round {r} {o}: {d['ms_per_step']:.3f} ms/step  parity rel-L2 {d['parity']['rel_l2']:.3e} max-abs {d['parity']['max_abs']:.3e}  "
                  f"clocks {d['clocks']['sm_mhz']}", flush=True)
        except Exception as e:
            print("failed", o, e, out.stderr[-1500:])
