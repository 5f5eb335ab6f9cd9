"""Ring FFT stages of ONE rank of an N-rank decomposition on one GPU (no exchange: the b input is random), to separate
layout / launch-tail effects from NVLink effects: python tools/fft_rank_bench.py <order> <lmax> <nranks> [reps]"""
import sys
import torch
sys.path.insert(0, ".")
import calclens_b200 as clb

order, lmax, nranks = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mode = sys.argv[5] if len(sys.argv) > 5 else "rr"      # m ownership: rr = round-robin (default), block = contiguous ranges
import numpy as np
from calclens_b200 import sht
ro, mo = sht.default_owners(order, lmax, nranks)
if mode == "block":
    per = -(-(lmax + 1) // nranks)
    mo = (np.arange(lmax + 1) // per).astype(np.int32)
plan = clb.HEALPixSHTPlan(order, lmax, None, nranks, 0, rp_owner=ro, m_owner=mo)
b = torch.randn(2 * max(plan.b_recv_total, 1), dtype=torch.float64, device="cuda")
m = torch.randn(plan.npix, device="cuda", dtype=torch.float32)
maps = torch.zeros((6, plan.npix), dtype=torch.float32, device="cuda")
g = torch.empty(2 * max(plan.g_send_total, 1), dtype=torch.float64, device="cuda")


def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ta = timeit(lambda: plan.ring_analysis(m, g))
ts = timeit(lambda: plan.ring_synthesis(b, maps))
print(mode, "rank 0 of %d: ring_analysis %.3f ms (x%d = %.2f), ring_synthesis %.3f ms (x%d = %.2f); %d local ring pairs" % (
    nranks, ta, nranks, ta * nranks, ts, nranks, ts * nranks, plan.nrp_loc))
