"""Parity margins at the benched depth (what the 2e-3 bar has left): prints every tensor's rel-L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G
for name, kw in (("12/24 64px x1", dict(weight_scale=1.0)), ("12/24 64px x3", dict(weight_scale=3.0))):
    r = G.rollout_case(2, 12, 24, 12, 64, 12, 64, 64, **kw)
    worst = sorted(r.items(), key=lambda kv: -kv[1])[:6]
    print(name, " ".join(f"{k}={v:.2e}" for k, v in worst))
