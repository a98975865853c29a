#!/usr/bin/env python
"""Build libsvt_b200.so from the csrc/ + include/ of a git revision into ab/<name>.so (git-ignored, travels with gpurun),
for same-box A/B runs with tools/ab_step.py.     python tools/build_variant.py <git-rev|WORKTREE> <name> [extra nvcc flags]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svt_speechbrain_b200 import build as B  # noqa: E402

rev, name, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
out_dir = os.path.join(ROOT, "ab")
os.makedirs(out_dir, exist_ok=True)
with tempfile.TemporaryDirectory() as tmp:
    if rev == "WORKTREE":
        subprocess.check_call(f"cd {ROOT} && tar -c svt_speechbrain_b200/csrc include | tar -x -C {tmp}", shell=True)
    else:
        subprocess.check_call(f"git -C {ROOT} archive {rev} svt_speechbrain_b200/csrc include | tar -x -C {tmp}", shell=True)
    csrc = os.path.join(tmp, "svt_speechbrain_b200", "csrc")
    objs, procs = [], []
    for src in B.SOURCES:
        if not os.path.exists(os.path.join(csrc, src)):
            continue
        obj = os.path.join(tmp, src + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([B._nvcc(), *B.NVCC_FLAGS, *extra, "-c", os.path.join(csrc, src), "-o", obj],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    assert all(p.wait() == 0 for p in procs)
    out = os.path.join(out_dir, name + ".so")
    subprocess.check_call([B._nvcc(), "-shared", "-cudart", "static", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    print(out)
