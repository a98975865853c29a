#!/usr/bin/env python
"""The audio-visual leg of bench.py (config 4, 4 clips per GPU) for two builds, alternating on one box.
    python tools/av_ab.py ab/A.so ab/B.so [rounds]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:3]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for r in range(rounds):
    for l in libs:
        env = dict(os.environ, SVT_B200_OPTIONS=l) if "=" in l else dict(os.environ, SVT_B200_LIB=os.path.abspath(l))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline",
                              "--no-parity", "--no-e2e"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(f"round {r} {os.path.basename(l)}: AV {d['aux']['av']['ms_per_step']:.3f} ms/step = {d['aux']['av']['audio_s_per_s']:.0f} audio-s/s, "
                  f"long-form {1e3 * d['aux']['longform']['wall_s']:.2f} ms, main step {d['ms_per_step']:.3f} ms", flush=True)
        except Exception as e:
            print("failed:", l, e, out.stderr[-1500:])
