"""Device time of the forward chain alone (plan.forward through the C ABI, CUDA events; no Python autograd, no weight
repack): persistent chain vs one launch per cell step, with and without CUDA-graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from satflow_b200 import ConvLSTM
from satflow_b200.plan import RolloutPlan

def run(B, tin, tout, hid, hw, training, env):
    for k, v in env.items():
        os.environ[k] = v
    torch.manual_seed(0)
    net = ConvLSTM(12, hid, 12).cuda()
    plan = RolloutPlan(B, hw, hw, 12, hid, 12, tin, tout, training=training, device=torch.device("cuda", 0))
    plan.set_weights(net.rollout_params())
    x = torch.randn(B, tin, 12, hw, hw, device="cuda")
    y = torch.empty(B, 12, tout, hw, hw, device="cuda")
    for _ in range(10):
        plan.forward(x, y)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    a.record()
    for _ in range(n):
        plan.forward(x, y)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    info = (plan.info("persistent_chain"), plan.info("graph_replays"))
    plan.close()
    for k in env:
        del os.environ[k]
    return us, info

for shape in ((2, 4, 4, 32, 64), (1, 12, 24, 64, 128), (4, 12, 24, 32, 64)):
    for training in (False, True):
        row = []
        for env in ({"CLSTM_PERSIST": "0", "CLSTM_GRAPH": "0"}, {"CLSTM_PERSIST": "0", "CLSTM_GRAPH": "1"},
                    {"CLSTM_PERSIST": "1", "CLSTM_GRAPH": "0"}, {"CLSTM_PERSIST": "1", "CLSTM_GRAPH": "1"}):
            us, info = run(*shape, training, env)
            row.append(f"persist={env['CLSTM_PERSIST']} graph={env['CLSTM_GRAPH']}: {us:8.1f} us")
        steps = 2 * (shape[1] + shape[2])
        print(f"B={shape[0]} {shape[1]}/{shape[2]} hid {shape[3]} {shape[4]}x{shape[4]} training={int(training)} ({steps} cell steps): " + " | ".join(row), flush=True)
