"""A/B of environment knobs on the inference rollout (BASELINE configs[1]) inside one process.
Usage: python tools/ab_infer.py "CLSTM_PAIR=0" "CLSTM_PAIR=1" ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import quick_bench as q

variants = sys.argv[1:] or ["CLSTM_PAIR=0", "CLSTM_PAIR=1"]
for rep in range(3):
    for v in variants:
        kv = dict(a.split("=") for a in v.split())
        for k, val in kv.items():
            os.environ[k] = val
        print(v, flush=True)
        q.run(16, 12, 24, 64, 256, False, iters=5)
        for k in kv:
            del os.environ[k]
