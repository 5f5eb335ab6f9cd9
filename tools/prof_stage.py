"""Run one SHT stage a few times (for ncu captures): python tools/prof_stage.py <order> <lmax> [reps]"""
import sys
import torch
sys.path.insert(0, ".")
import calclens_b200 as clb

order, lmax = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
nshell = int(sys.argv[4]) if len(sys.argv) > 4 else 1
import os
from calclens_b200 import _lib
if os.environ.get("CLB_SYN_R"): _lib.load().clb_set_tuning(0, int(os.environ["CLB_SYN_R"]))
if os.environ.get("CLB_ANA_R"): _lib.load().clb_set_tuning(1, int(os.environ["CLB_ANA_R"]))
plan = clb.HEALPixSHTPlan(order, lmax)
m = torch.randn(plan.npix, device="cuda", dtype=torch.float32)
g = torch.empty(2 * nshell * plan.g_send_total, dtype=torch.float64, device="cuda")
for _ in range(reps):
    for s in range(nshell):
        plan.ring_analysis(m, g[2 * s * plan.g_send_total:])
    are, aim = plan.legendre_analysis(g, poisson_filter=True, nshell=nshell)
    b = plan.legendre_synthesis(are, aim, nshell=nshell)
    maps = plan.ring_synthesis(b)
torch.cuda.synchronize()
print("done")
