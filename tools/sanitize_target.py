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
# round 2: the persistent forward chain + CUDA-graph replay run on all of the above (every shape fits 148 CTAs);
# the per-step kernels, both fused-dgrad generations and the native cell stepper on a second pass
import torch
from satflow_b200 import ConvLSTMCell
for env in ({"CLSTM_PERSIST": "0", "CLSTM_GRAPH": "0"}, {"CLSTM_FUSE_WORKERS": "0"}, {"CLSTM_HYBRID": "50"}):
    os.environ.update(env)
    r = G.rollout_case(1, 2, 3, 12, 64, 5, 4, 256, states=False)
    print(env, max(r.values()))
    for k in env:
        del os.environ[k]
# session 2: the CTA-pair cell step (cta_group::2; needs two waves of tiles: 384 here) and, in every rollout above, the
# fused dgrad with the head's dgrad as a second K segment and c' recomputed from the gates
r = G.rollout_case(1, 1, 2, 12, 64, 12, 128, 384, states=False)
print("pair", max(r.values()))
r = G.cell_unrolled_case(1, 12, 16, 6, 20, 3)
print("unrolled cell", max(r.values()))
cell = ConvLSTMCell(12, 32, (3, 3), True).cuda()
st = cell.native(1, (6, 40)).step(torch.randn(1, 12, 6, 40, device="cuda")).step()
print("native", float(st.state()[0].abs().max()))
