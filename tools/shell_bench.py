"""Time the Legendre kernels with one and two shells per pass (CUDA events), for a list of tuning settings.
   python tools/shell_bench.py <order> <lmax> [ana2=R,R,...] [syn2=R,R,...] [rows=n,n,...] [reps]"""
import sys
import torch
sys.path.insert(0, ".")
import calclens_b200 as clb
from calclens_b200 import _lib

order, lmax = int(sys.argv[1]), int(sys.argv[2])
ana2 = [8]; syn2 = [4]; rows = [0]; reps = 3; pipes = [1]
for a in sys.argv[3:]:
    if a.startswith("ana2="): ana2 = [int(x) for x in a[5:].split(",")]
    elif a.startswith("syn2="): syn2 = [int(x) for x in a[5:].split(",")]
    elif a.startswith("rows="): rows = [int(x) for x in a[5:].split(",")]
    elif a.startswith("pipe="): pipes = [int(x) for x in a[5:].split(",")]
    else: reps = int(a)
L = _lib.load()
plan = clb.HEALPixSHTPlan(order, lmax)
gt, n, bt = plan.g_send_total, plan.Nlm, plan.b_send_total
g = torch.empty(4 * gt, dtype=torch.float64, device="cuda")
for s in range(2):
    m = torch.randn(plan.npix, device="cuda", dtype=torch.float32)
    plan.ring_analysis(m, g[2 * s * gt:])
are = torch.empty(2 * n, dtype=torch.float64, device="cuda"); aim = torch.empty_like(are)
b = torch.empty(4 * bt, dtype=torch.float64, device="cuda")


def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for pipe in pipes:
    L.clb_set_tuning(12, 2 * pipe)
    for r in rows:
        L.clb_set_tuning(5, r)
        t = timeit(lambda: plan.legendre_analysis(g, are, aim, poisson_filter=True))
        print("analysis  1 shell  pipe=%d rows=%d           %8.3f ms per shell" % (pipe, r, t))
        ref = (are[:n].clone(), aim[:n].clone())
        for R in ana2:
            L.clb_set_tuning(10, R)
            t = timeit(lambda: plan.legendre_analysis(g, are, aim, poisson_filter=True, nshell=2))
            print("analysis  2 shells pipe=%d rows=%d R=%d       %8.3f ms per shell   shell 0 identical to one-shell: %s" % (
                pipe, r, R, t / 2, bool(torch.equal(are[:n], ref[0]) and torch.equal(aim[:n], ref[1]))))
L.clb_set_tuning(5, 0); L.clb_set_tuning(10, 8); L.clb_set_tuning(12, 1)
plan.legendre_analysis(g, are, aim, poisson_filter=True, nshell=2)
t = timeit(lambda: plan.legendre_synthesis(are, aim, b))
print("synthesis 1 shell                    %8.3f ms per shell" % t)
bref = b[:2 * bt].clone()
for R in syn2:
    L.clb_set_tuning(9, R)
    t = timeit(lambda: plan.legendre_synthesis(are, aim, b, nshell=2))
    print("synthesis 2 shells R=%d              %8.3f ms per shell   shell 0 identical to one-shell: %s" % (R, t / 2, bool(torch.equal(b[:2 * bt], bref))))
