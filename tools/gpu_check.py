"""Quick stage-by-stage GPU-vs-oracle check (development aid; the real parity tests live in tests/)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import calclens_b200 as clb
from calclens_b200 import _lib, rays as crays
from oracle import ref

L = _lib.load()
print("devices", L.clb_device_count())
dev = torch.device("cuda:0")
rng = np.random.default_rng(7)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300))


# --- A. indexing
for order in (0, 1, 3, 6):
    npix = 12 << (2 * order)
    n = min(npix, 20000)
    pix = rng.integers(0, npix, n) if npix > n else np.arange(npix)
    R = ref.lib()
    r2n = np.array([R.ring2nest(int(p), order) for p in pix]); n2r = np.array([R.nest2ring(int(p), order) for p in pix])
    n2p = np.array([R.nest2peano(int(p), order) for p in pix])
    th = np.arccos(rng.uniform(-1, 1, n)); ph = rng.uniform(0, 2 * np.pi, n)
    a2n = np.array([R.ang2nest(float(t), float(p), order) for t, p in zip(th, ph)])
    tin = torch.from_numpy(pix.astype(np.int64)).to(dev); tth = torch.from_numpy(th).to(dev); tph = torch.from_numpy(ph).to(dev)
    out = torch.empty(n, dtype=torch.int64, device=dev)
    res = []
    for what, want in ((0, r2n), (1, n2r), (2, a2n), (3, n2p)):
        L.clb_healpix_index_dev(what, order, n, tin.data_ptr(), tth.data_ptr(), tph.data_ptr(), out.data_ptr(), None)
        torch.cuda.synchronize()
        res.append(int((out.cpu().numpy() != want).sum()))
    print("order", order, "index mismatches ring2nest/nest2ring/ang2nest/nest2peano:", res)

# --- B. rays
order = 5; bo = 2
npix = 12 << (2 * order)
maps = (rng.normal(size=(6, npix)) * np.array([1, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2])[:, None]).astype(np.float32)
rays0 = ref.init_rays(6, 15.0)
rays0["n"] += rng.normal(size=rays0["n"].shape) * 0.01
ra = rays0.copy(); rb = rays0.copy()
ref.shearinterp(order, bo, maps, ra)
crays.shearinterp_rays(maps, order, rb)
for f in ("phi", "alpha", "U"):
    print("interp", f, "rel", rel(rb[f], ra[f]), "maxabs", np.abs(rb[f] - ra[f]).max())
ref.rayprop(ra, 45.0, 15.0, 0.0)
crays.rayprop_sphere(45.0, 15.0, 0.0, rb)
for f in ("n", "beta", "A", "Aprev"):
    print("prop1", f, "rel", rel(rb[f], ra[f]), "maxabs", np.abs(rb[f] - ra[f]).max())
ra["alpha"] = 0; ra["U"] = 0; rb["alpha"] = 0; rb["U"] = 0
ref.shearinterp(order, bo, maps, ra); crays.shearinterp_rays(maps, order, rb)
ref.rayprop(ra, 75.0, 45.0, 15.0); crays.rayprop_sphere(75.0, 45.0, 15.0, rb)
for f in ("n", "beta", "A", "Aprev", "alpha", "U"):
    print("prop2", f, "rel", rel(rb[f], ra[f]), "maxabs", np.abs(rb[f] - ra[f]).max())

# --- C/D. SHT
for order, lmax, useW in ((2, 8, False), (4, 32, False), (4, 47, True), (6, 128, True), (7, 256, False), (8, 512, True)):
    nside = 1 << order; npix = 12 * nside * nside
    w = None
    try:
        w = clb.read_ring_weights("/root/reference/healpix_weights", order) if useW else None
    except Exception:
        w = rng.normal(size=2 * nside) * 0.01 if useW else None
    m = rng.lognormal(size=npix).astype(np.float32)
    m -= m.mean()
    t = time.time(); are, aim = ref.map2alm(order, lmax, m, w); tref = time.time() - t
    plan = clb.HEALPixSHTPlan(order, lmax, ring_weights=w)
    t = time.time(); gre, gim = clb.map2alm_mpi(m, plan); tg = time.time() - t
    e = np.sqrt((((gre - are) ** 2 + (gim - aim) ** 2).sum()) / ((are ** 2 + aim ** 2).sum()))
    print("order %d lmax %d w %s: map2alm rel L2 %.3e (ref %.2fs gpu %.2fs) max|d| %.2e" % (order, lmax, useW, e, tref, tg, max(np.abs(gre - are).max(), np.abs(gim - aim).max())))
    fre, fim = ref.poisson_filter(lmax, are, aim)
    t = time.time(); mref = ref.alm2allmaps(order, lmax, fre, fim); tref = time.time() - t
    t = time.time(); mg = clb.alm2allmaps_mpi(fre, fim, plan); tg = time.time() - t
    for k in range(6):
        neq = int((mg[k] != mref[k]).sum())
        print("    field %d rel %.3e  differing floats %d / %d  maxabs %.3e (scale %.3e)" % (k, rel(mg[k], mref[k]), neq, npix, np.abs(mg[k] - mref[k]).max(), np.abs(mref[k]).max()))
    print("    (ref %.2fs gpu %.2fs)" % (tref, tg))
    plan.destroy()
print("launches", L.clb_launch_count())

# --- timing of the device-resident stages
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for order, lmax in ((9, 1024), (10, 2048), (11, 4096), (12, 8192)):
    if len(sys.argv) > 1 and order > int(sys.argv[1]):
        break
    nside = 1 << order; npix = 12 * nside * nside
    t = time.time(); plan = clb.HEALPixSHTPlan(order, lmax); torch.cuda.synchronize(); tplan = time.time() - t
    m = torch.randn(npix, device=dev, dtype=torch.float32)
    g = plan.ring_analysis(m)
    are, aim = plan.legendre_analysis(g, poisson_filter=True)
    b = plan.legendre_synthesis(are, aim)
    maps = plan.ring_synthesis(b)
    nrays = npix
    rays = torch.zeros(nrays * 176, dtype=torch.uint8, device=dev)
    t_fa = timeit(lambda: plan.ring_analysis(m, g))
    t_la = timeit(lambda: plan.legendre_analysis(g, are, aim, True))
    t_ls = timeit(lambda: plan.legendre_synthesis(are, aim, b))
    t_fs = timeit(lambda: plan.ring_synthesis(b, maps))
    print("order %d lmax %d: plan %.2fs | fft_ana %.2f ms  leg_ana %.2f ms  leg_syn %.2f ms  fft_syn %.2f ms | mem %.1f GB" % (
        order, lmax, tplan, t_fa, t_la, t_ls, t_fs, torch.cuda.max_memory_allocated() / 1e9), flush=True)
    plan.destroy(); del m, g, are, aim, b, maps, rays
    torch.cuda.empty_cache()
