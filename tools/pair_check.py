"""Quick parity check of the CTA-pair cell step at sizes that select it (used while bringing cellstep_pair.cuh up)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G
r = G.rollout_case(2, 2, 2, 12, 64, 12, 256, 256)
worst = sorted(r.items(), key=lambda kv: -kv[1])[:5]
print("PAIR=%s" % os.environ.get("CLSTM_PAIR"), " ".join(f"{k}={v:.2e}" for k, v in worst), flush=True)
assert max(r.values()) <= 2e-3
r = G.rollout_case(1, 3, 3, 12, 64, 12, 128, 384, backward=False)
print("fwd-only 128x384", max(r.values()), flush=True)
assert max(r.values()) <= 2e-3
