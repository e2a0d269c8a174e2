"""Small fwd+bwd through every default kernel (incl. the 128-pixel-row paths) for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G
r = G.rollout_case(1, 2, 2, 12, 64, 5, 4, 256, states=False)
print("W256", max(r.values()))
r = G.rollout_case(2, 2, 2, 12, 16, 3, 9, 10, states=False)
print("small", max(r.values()))
r = G.rollout_case(1, 1, 2, 12, 8, 12, 34, 300, states=False)  # head strips + row bands, fused dgrad on ragged tiles
print("W300", max(r.values()))
r = G.cell_case(1, 12, 32, 6, 40, 3, 3)
print("cell", max(r.values()))
