"""compute-sanitizer target for the LARGE-plan default path (384 tiles): CTA-pair cell step, 16-bit c / dc / dh, fused backward
chain with the head's dgrad as a K segment, on a 2-layer and a 3-layer stack (forward + backward, checked against the oracle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G
r = G.rollout_case(1, 2, 2, 12, 64, 12, 128, 384)
print("large 2-layer", max(r.values()))
r = G.rollout_case(1, 1, 2, 12, 64, 12, 128, 384, n_layers=3, states=False)
print("large 3-layer", max(r.values()))
