"""Parity margins of one rollout case under different knob settings: python tools/margin_case.py B tin tout cin hid cout H W layers k"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G
a = [int(v) for v in sys.argv[1:11]]
for env in ({}, {"CLSTM_C16": "0"}, {"CLSTM_C16": "0", "CLSTM_STATE16": "0"},
            {"CLSTM_C16": "0", "CLSTM_STATE16": "0", "CLSTM_RECOMP_C": "0", "CLSTM_HEAD_FUSE": "0", "CLSTM_PAIR": "0"}):
    os.environ.update(env)
    r = G.rollout_case(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], n_layers=a[8], k=a[9], seed=3)
    worst = sorted(r.items(), key=lambda kv: -kv[1])[:4]
    print(env, " ".join(f"{k}={v:.2e}" for k, v in worst), flush=True)
    for k in env:
        del os.environ[k]
