#!/usr/bin/env python
"""A/B of two builds of libsvt_b200.so on ONE box: the step is power-limited (sw_power_cap), and box-to-box spread (+-2 %) is
larger than most single optimisations, so builds are compared by running bench.py's main timed region alternately
(A B A B ...) in fresh processes on the same GPU.
    python tools/ab_step.py path/to/libA.so path/to/libB.so [rounds] [steps]
An argument of the form name=value[,name=value] instead of a path runs the in-tree build with those svt_set_option
switches (SVT_B200_OPTIONS), e.g.   python tools/ab_step.py rowln_fuse=0 rowln_fuse=1"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:3]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
steps = sys.argv[4] if len(sys.argv) > 4 else "20"
res = {l: [] for l in libs}
for r in range(rounds):
    for l in libs:
        env = dict(os.environ, SVT_B200_OPTIONS=l) if "=" in l else dict(os.environ, SVT_B200_LIB=os.path.abspath(l))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "5", "--no-aux",
                              "--no-cpu-baseline", "--no-parity", "--no-e2e"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            res[l].append(d["ms_per_step"])
            print(f"round {r} {os.path.basename(l)}: {d['ms_per_step']:.3f} ms/step, clocks {d['clocks']['sm_mhz']} MHz, "
                  f"ffn1 {d['roofline']['us_per_launch']:.1f} us", flush=True)
        except Exception as e:
            print("failed:", l, e, out.stderr[-2000:])
for l in libs:
    v = sorted(res[l])
    if v:
        print(f"{os.path.basename(l)}: median {v[len(v) // 2]:.3f} ms/step  all {['%.3f' % x for x in res[l]]}")
