"""A/B of two builds of the library on the full training step: alternates subprocesses with CLSTM_LIB set.
Usage: python tools/ab_lib.py satflow_b200/libclstm_old.so [reps]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
old = os.path.abspath(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import quick_bench as q; "
        "q.run(16, 12, 24, 64, 256, True, iters=3)" % (ROOT, os.path.join(ROOT, "tools")))
for _ in range(reps):
    for tag, lib in (("OLD", old), ("NEW", "")):
        env = dict(os.environ)
        if lib:
            env["CLSTM_LIB"] = lib
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout
        print(tag, out.strip().split("fp16:")[-1], flush=True)
