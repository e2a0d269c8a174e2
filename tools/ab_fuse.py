"""A/B of environment knobs on the full training step, inside one process (same box, same clocks).
Usage: python tools/ab_fuse.py "CLSTM_FUSE_GATE=0" "CLSTM_FUSE_GATE=1 CLSTM_FUSE_PF=2" ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import quick_bench as q

variants = sys.argv[1:] or ["CLSTM_FUSE_GATE=0", "CLSTM_FUSE_GATE=1"]
for rep in range(2):
    for v in variants:
        kv = dict(a.split("=") for a in v.split())
        for k, val in kv.items():
            os.environ[k] = val
        print(v, flush=True)
        q.run(16, 12, 24, 64, 256, True, iters=3)
        for k in kv:
            del os.environ[k]
