"""compute-sanitizer target for the CTA-pair cell step only (384 tiles: two waves)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G
r = G.rollout_case(1, 1, 2, 12, 64, 12, 128, 384, states=False)
print("pair", max(r.values()))
