"""Time the four SHT stages separately with CUDA events, for a list of tuning settings.
   python tools/stage_bench.py <order> <lmax> [syn=R,R,...] [ana=R,R,...] [reps]"""
import sys
import torch
sys.path.insert(0, ".")
import calclens_b200 as clb
from calclens_b200 import _lib

order, lmax = int(sys.argv[1]), int(sys.argv[2])
syn = [4]; ana = [8]; reps = 3
for a in sys.argv[3:]:
    if a.startswith("syn="): syn = [int(x) for x in a[4:].split(",")]
    elif a.startswith("ana="): ana = [int(x) for x in a[4:].split(",")]
    elif a.startswith("warps="): _lib.load().clb_set_tuning(3, int(a[6:]))
    elif a.startswith("fft="): _lib.load().clb_set_tuning(2, int(a[4:]))
    elif a.startswith("scratch="): _lib.load().clb_set_tuning(4, int(a[8:]))
    elif a.startswith("rows="): _lib.load().clb_set_tuning(5, int(a[5:]))
    elif a.startswith("streams="): _lib.load().clb_set_tuning(7, int(a[8:]))
    elif a.startswith("pipe="): _lib.load().clb_set_tuning(12, int(a[5:]))
    elif a.startswith("fg="): _lib.load().clb_set_tuning(6, int(a[3:]))
    elif a.startswith("dbg="): _lib.load().clb_set_tuning(8, int(a[4:]))
    else: reps = int(a)
L = _lib.load()
plan = clb.HEALPixSHTPlan(order, lmax)
m = torch.randn(plan.npix, device="cuda", dtype=torch.float32)


def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


g_buf = torch.empty(2 * max(plan.g_send_total, 1), dtype=torch.float64, device="cuda")
are_b = torch.empty(max(plan.Nlm, 1), dtype=torch.float64, device="cuda"); aim_b = torch.empty_like(are_b)
b_buf = torch.empty(2 * max(plan.b_send_total, 1), dtype=torch.float64, device="cuda")
maps_b = torch.zeros((6, plan.npix), dtype=torch.float32, device="cuda")
t, g = timeit(lambda: plan.ring_analysis(m, g_buf))
print("ring_analysis        %8.3f ms" % t)
ref = None
for r in ana:
    L.clb_set_tuning(1, r)
    t, (are, aim) = timeit(lambda: plan.legendre_analysis(g, are_b, aim_b, poisson_filter=True))
    if ref is None:
        ref = (are.clone(), aim.clone())
    err = float(((are - ref[0]).norm() ** 2 + (aim - ref[1]).norm() ** 2).sqrt() / (ref[0].norm() ** 2 + ref[1].norm() ** 2).sqrt())
    print("legendre_analysis R=%d %8.3f ms   rel diff to first %.2e" % (r, t, err))
bref = None
for r in syn:
    L.clb_set_tuning(0, r)
    t, b = timeit(lambda: plan.legendre_synthesis(ref[0], ref[1], b_buf))
    if bref is None:
        bref = b.clone()
    print("legendre_synthesis R=%d %8.3f ms  identical to first: %s" % (r, t, bool(torch.equal(b, bref))))
t, maps = timeit(lambda: plan.ring_synthesis(bref, maps_b))
print("ring_synthesis       %8.3f ms" % t)
# ray stage on the maps just synthesised (scaled to lensing-like amplitudes)
import ctypes as C
maps = maps_b * (1e-3 / max(float(maps_b[3].abs().max()), 1e-30))
nrays = plan.npix
rays = torch.empty(nrays * 176, dtype=torch.uint8, device="cuda")
L.clb_ray_init_dev(rays.data_ptr(), nrays, 0, order, 15.0, None)
ptrs = (C.c_void_p * 6)(*[maps[k].data_ptr() for k in range(6)])
torch.cuda.synchronize()
def raystep():
    L.clb_ray_init_dev(rays.data_ptr(), nrays, 0, order, 15.0, None)
    L.clb_ray_step_dev(rays.data_ptr(), nrays, ptrs, order, 45.0, 15.0, 0.0, 7, None)
t_both, _ = timeit(raystep)
t_init, _ = timeit(lambda: L.clb_ray_init_dev(rays.data_ptr(), nrays, 0, order, 15.0, None))
print("ray_step (interp+prop) %8.3f ms  (%d rays; init %.3f ms subtracted)" % (t_both - t_init, nrays, t_init))
