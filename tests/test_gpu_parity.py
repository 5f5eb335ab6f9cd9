"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs, against the committed golden vectors, and -- at the benchmark sizes, where the CPU oracle is too
slow -- through size-independent properties.

Tolerances (BASELINE.json north_star): pixel/ring indexing bit-exact; alm <= 1e-10 relative L2 (FP64); ray
convergence, shear, deflection <= 1e-8 relative.  The six float32 derivative maps are compared bit for bit: the CUDA
path rounds to float exactly where the reference does, so at most a handful of pixels may differ by one float ulp
(FP64 round-off before a rounding boundary); the bound asserted is <= 1e-5 of the pixels and <= 2e-7 relative L2.
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALM_TOL = 1e-10
RAY_TOL = 1e-8


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def alm_err(gre, gim, are, aim):
    return float(np.sqrt((((gre - are) ** 2 + (gim - aim) ** 2).sum()) / max((are ** 2 + aim ** 2).sum(), 1e-300)))


def assert_maps_match(mg, mo, what):
    """float maps: a pixel counts as different when it differs by more than one float ulp of its own value plus
    1e-9 of the field's scale (pixels that vanish by symmetry hold pure FP64 round-off, ~1e-16 of the scale, on
    both sides and cannot agree bit for bit)"""
    for k in range(6):
        a = mg[k].astype(np.float64); b = mo[k].astype(np.float64)
        scale = np.abs(b).max()
        bad = np.abs(a - b) > 1.2e-7 * np.maximum(np.abs(a), np.abs(b)) + 1e-9 * scale
        ndiff = int(bad.sum())
        assert ndiff <= max(2, 1e-5 * b.size), "%s field %d: %d of %d floats differ" % (what, k, ndiff, b.size)
        assert rel_l2(a, b) <= 2e-7, "%s field %d rel L2 %g" % (what, k, rel_l2(a, b))


def assert_rays_match(rg, ro, fields=("n", "beta", "A", "Aprev", "alpha", "U", "phi")):
    for f in fields:
        scale = max(np.abs(ro[f]).max(), 1e-300)
        d = np.abs(rg[f] - ro[f]).max()
        assert d <= RAY_TOL * scale, "ray field %s: max abs diff %g (scale %g)" % (f, d, scale)


@pytest.fixture(scope="module")
def clb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import calclens_b200
    from calclens_b200 import _lib
    assert _lib.load().clb_device_count() >= 1
    return calclens_b200


@pytest.fixture(scope="module")
def weights():
    return np.load(os.path.join(GOLD, "ring_weights.npz"))


# ----------------------------------------------------------------------------------------------------------------
# indexing: bit-exact
# ----------------------------------------------------------------------------------------------------------------
def _index_dev(what, order, pix=None, theta=None, phi=None):
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    n = len(pix) if pix is not None else len(theta)
    dev = torch.device("cuda")
    tin = torch.from_numpy(np.asarray(pix if pix is not None else np.zeros(n), dtype=np.int64)).to(dev)
    tth = torch.from_numpy(np.asarray(theta if theta is not None else np.zeros(n), dtype=np.float64)).to(dev)
    tph = torch.from_numpy(np.asarray(phi if phi is not None else np.zeros(n), dtype=np.float64)).to(dev)
    out = torch.empty(n, dtype=torch.int64, device=dev)
    L.clb_healpix_index_dev(what, order, n, tin.data_ptr(), tth.data_ptr(), tph.data_ptr(), out.data_ptr(), None)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_indexing_golden_bit_exact(clb):
    g = np.load(os.path.join(GOLD, "healpix_index.npz"))
    for order in (0, 1, 2, 3, 4, 5):
        p = np.arange(12 << (2 * order))
        assert np.array_equal(_index_dev(0, order, p), g["ring2nest_o%d" % order])
        assert np.array_equal(_index_dev(1, order, p), g["nest2ring_o%d" % order])
        assert np.array_equal(_index_dev(3, order, p), g["nest2peano_o%d" % order])
    for order in (0, 3, 8, 13):
        assert np.array_equal(_index_dev(2, order, theta=g["theta"], phi=g["phi"]), g["ang2nest_o%d" % order])


def test_indexing_vs_oracle_large(clb, oracle):
    rng = np.random.default_rng(5)
    for order in (7, 10, 12):
        npix = 12 << (2 * order)
        pix = rng.integers(0, npix, 4000)
        # ring boundaries: first/last pixel of caps and of the equatorial belt
        nside = 1 << order
        ncap = 2 * nside * (nside - 1)
        pix[:8] = [0, 3, 4, ncap - 1, ncap, npix - ncap - 1, npix - ncap, npix - 1]
        assert np.array_equal(_index_dev(0, order, pix), np.array([oracle.ring2nest(p, order) for p in pix]))
        assert np.array_equal(_index_dev(1, order, pix), np.array([oracle.nest2ring(p, order) for p in pix]))
        # round trip on the device
        assert np.array_equal(_index_dev(1, order, _index_dev(0, order, pix)), pix)
        th = np.arccos(rng.uniform(-1, 1, 4000)); ph = rng.uniform(0, 2 * np.pi, 4000)
        assert np.array_equal(_index_dev(2, order, theta=th, phi=ph), np.array([oracle.ang2nest(t, p, order) for t, p in zip(th, ph)]))


def test_interpolation_stencil_vs_oracle(clb, oracle):
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(6)
    for order in (1, 4, 10):
        n = 3000
        v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
        v[:4] = [[1e-4, 0, 1], [0, 1e-4, -1], [1, 0, 0], [0.6, 0.0, 0.8]]     # near both poles, equator
        v *= rng.uniform(10, 1000, n)[:, None]
        tv = torch.from_numpy(v.copy()).cuda()
        pix = torch.empty((n, 4), dtype=torch.int64, device="cuda"); wgt = torch.empty((n, 4), dtype=torch.float64, device="cuda")
        L.clb_healpix_interpol_dev(order, n, tv.data_ptr(), pix.data_ptr(), wgt.data_ptr(), None)
        torch.cuda.synchronize()
        pix = pix.cpu().numpy(); wgt = wgt.cpu().numpy()
        for i in range(n):
            r = np.linalg.norm(v[i])
            theta = np.arccos(v[i, 2] / r); phi = np.arctan2(v[i, 1], v[i, 0]) % (2 * np.pi)
            p, w = oracle.get_interpol(theta, phi, order)
            assert list(pix[i]) == p, (order, i)
            assert np.abs(wgt[i] - np.array(w)).max() < 1e-9


# ----------------------------------------------------------------------------------------------------------------
# SHT
# ----------------------------------------------------------------------------------------------------------------
def test_sht_golden(clb, weights):
    g = np.load(os.path.join(GOLD, "sht.npz"))
    for tag in "abcd":
        order, lmax = int(g[tag + "_order"]), int(g[tag + "_lmax"])
        w = weights["n%05d" % (1 << order)] if int(g[tag + "_weights"]) else None
        plan = clb.HEALPixSHTPlan(order, lmax, ring_weights=w)
        are, aim = clb.map2alm_mpi(g[tag + "_map"], plan)
        assert alm_err(are, aim, g[tag + "_alm_re"], g[tag + "_alm_im"]) <= ALM_TOL
        fre, fim = clb.map2alm_mpi(g[tag + "_map"], plan, poisson_filter=True)
        assert alm_err(fre, fim, g[tag + "_falm_re"], g[tag + "_falm_im"]) <= ALM_TOL
        maps = clb.alm2allmaps_mpi(g[tag + "_falm_re"], g[tag + "_falm_im"], plan)
        assert_maps_match(maps, g[tag + "_maps"], "golden " + tag)
        plan.destroy()


@pytest.mark.parametrize("order,lmax,use_w", [(0, 2, False), (1, 5, False), (2, 8, False), (4, 32, True), (4, 47, True),
                                              (5, 40, False), (6, 128, True), (6, 191, False), (7, 256, True), (8, 512, True)])
def test_sht_vs_oracle(clb, oracle, weights, order, lmax, use_w):
    """lmax = 3 Nside - 1 is the reference's own default (aliasing on every polar ring), lmax = 2 Nside the
    BASELINE configs', the odd ones exercise ragged block/tile tails."""
    rng = np.random.default_rng(100 + order)
    nside = 1 << order; npix = 12 * nside * nside
    w = weights["n%05d" % nside] if use_w else None
    m = ((8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32) * np.float32(3e-4) - np.float32(8 * np.exp(0.125) * 3e-4)).astype(np.float32)
    are, aim = oracle.map2alm(order, lmax, m, w)
    plan = clb.HEALPixSHTPlan(order, lmax, ring_weights=w)
    gre, gim = clb.map2alm_mpi(m, plan)
    assert alm_err(gre, gim, are, aim) <= ALM_TOL
    fre, fim = oracle.poisson_filter(lmax, are, aim)
    mo = oracle.alm2allmaps(order, lmax, fre, fim)
    mg = clb.alm2allmaps_mpi(fre, fim, plan)
    assert_maps_match(mg, mo, "order %d lmax %d" % (order, lmax))
    plan.destroy()


def test_sht_edge_inputs(clb, oracle):
    order, lmax = 4, 32
    npix = 12 << (2 * order)
    plan = clb.HEALPixSHTPlan(order, lmax)
    # empty map -> exactly zero alm and maps
    are, aim = clb.map2alm_mpi(np.zeros(npix, dtype=np.float32), plan)
    assert not are.any() and not aim.any()
    assert not clb.alm2allmaps_mpi(are, aim, plan).any()
    # point mass (single non-zero pixel) at a pole pixel, an equator pixel and a cap boundary pixel
    for pix in (0, npix // 2, 2 * 16 * 15, npix - 1):
        m = np.zeros(npix, dtype=np.float32); m[pix] = 1.0
        a = oracle.map2alm(order, lmax, m); b = clb.map2alm_mpi(m, plan)
        assert alm_err(b[0], b[1], a[0], a[1]) <= ALM_TOL
        f = oracle.poisson_filter(lmax, *a)
        assert_maps_match(clb.alm2allmaps_mpi(f[0], f[1], plan), oracle.alm2allmaps(order, lmax, *f), "point mass %d" % pix)
    # constant map: only a_00 (removed by the filter)
    a = clb.map2alm_mpi(np.full(npix, 2.5, dtype=np.float32), plan, poisson_filter=True)
    assert a[0][0] == 0.0 and a[1][0] == 0.0
    plan.destroy()


def test_mapvec_entry_points(clb, oracle):
    """the reference's padded ring-pair buffer layout (healpix_shtrans.c:90-118) through clb_*_mapvec"""
    from calclens_b200 import _lib
    L = _lib.load()
    order, lmax = 3, 16
    nside = 1 << order; npix = 12 * nside * nside
    rng = np.random.default_rng(9)
    m = rng.normal(size=npix).astype(np.float32)
    # build the layout exactly as healpixsht_plan does for one rank
    ns, ss, off = [], [], 0
    for r in range(1, 2 * nside + 1):
        n = 4 * min(r, nside)
        ns.append(off); off += n // 2 + 1
        if r != 2 * nside:
            ss.append(off); off += n // 2 + 1
        else:
            ss.append(-1)
    assert off == npix // 2 + 4 * nside - 1          # Nmapvec, SURVEY.md section 4
    ns = np.array(ns, dtype=np.int64); ss = np.array(ss, dtype=np.int64)
    mapvec = np.zeros(2 * off, dtype=np.float32)
    for i, r in enumerate(range(1, 2 * nside + 1)):
        n = 4 * min(r, nside)
        start = 2 * r * (r - 1) if r < nside else 2 * nside * (nside - 1) + (r - nside) * 4 * nside
        mapvec[2 * ns[i]:2 * ns[i] + n] = m[start:start + n]
        if ss[i] >= 0:
            mapvec[2 * ss[i]:2 * ss[i] + n] = m[npix - start - n:npix - start]
    plan = clb.HEALPixSHTPlan(order, lmax)
    are = np.empty(plan.Nlm); aim = np.empty(plan.Nlm)
    L.clb_map2alm_mapvec(plan._h, mapvec.ctypes.data, ns.ctypes.data, ss.ctypes.data, are.ctypes.data, aim.ctypes.data)
    ore, oim = oracle.map2alm(order, lmax, m)
    assert alm_err(are, aim, ore, oim) <= ALM_TOL
    bufs = [np.zeros(2 * off, dtype=np.float32) for _ in range(6)]
    ptrs = (C.c_void_p * 6)(*[b.ctypes.data for b in bufs])
    L.clb_alm2allmaps_mapvec(plan._h, ore.ctypes.data, oim.ctypes.data, ptrs, ns.ctypes.data, ss.ctypes.data)
    mo = oracle.alm2allmaps(order, lmax, ore, oim)
    for k in range(6):
        for i, r in enumerate(range(1, 2 * nside + 1)):
            n = 4 * min(r, nside)
            start = 2 * r * (r - 1) if r < nside else 2 * nside * (nside - 1) + (r - nside) * 4 * nside
            assert np.array_equal(bufs[k][2 * ns[i]:2 * ns[i] + n], mo[k][start:start + n])
    plan.destroy()


def test_sht_properties_at_benchmark_size(clb):
    """Nside=1024, lmax=2048 (BASELINE configs[1]) without a CPU oracle: linearity, the Laplacian identity
    grad_tt + grad_pp = -kappa-like source, and analysis(synthesis(alm)) = alm to quadrature accuracy."""
    import torch
    order, lmax = 10, 2048
    plan = clb.HEALPixSHTPlan(order, lmax)
    dev = plan.device
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    ls = torch.cat([torch.arange(m, lmax + 1, device=dev, dtype=torch.float64) for m in range(lmax + 1)])
    amp = (ls + 10.0) ** -1.1
    are = torch.randn(plan.Nlm, generator=gen, device=dev, dtype=torch.float64) * amp
    aim = torch.randn(plan.Nlm, generator=gen, device=dev, dtype=torch.float64) * amp
    aim[:lmax + 1] = 0.0
    are[0] = 0.0
    maps = plan.ring_synthesis(plan.legendre_synthesis(are, aim)).clone()
    # Laplacian identity: maps[3] + maps[5] is the synthesis of -l(l+1) a_lm
    lap = plan.ring_synthesis(plan.legendre_synthesis(-ls * (ls + 1) * are, -ls * (ls + 1) * aim))[0]
    num = (maps[3].double() + maps[5].double() - lap.double()).pow(2).sum().sqrt()
    assert float(num / lap.double().pow(2).sum().sqrt()) < 2e-5
    # linearity of synthesis (float32 output): S(2a) == 2 S(a) exactly, S(a+b) ~ S(a)+S(b)
    maps2 = plan.ring_synthesis(plan.legendre_synthesis(2 * are, 2 * aim))
    assert torch.equal(maps2[0], 2 * maps[0])
    # round trip through analysis (no ring weights: HEALPix quadrature error ~1e-3 at lmax = 2 Nside)
    bre, bim = plan.legendre_analysis(plan.ring_analysis(maps[0].contiguous()))
    err = float(((bre - are).pow(2) + (bim - aim).pow(2)).sum().sqrt() / (are.pow(2) + aim.pow(2)).sum().sqrt())
    assert err < 5e-3
    plan.destroy()


# ----------------------------------------------------------------------------------------------------------------
# rays
# ----------------------------------------------------------------------------------------------------------------
def test_rays_golden(clb):
    g = np.load(os.path.join(GOLD, "rays.npz"))
    order = int(g["order"])
    rays = g["rays0"].view(clb.RAY_DTYPE).copy()
    clb.shearinterp_rays(g["maps"], order, rays)
    assert_rays_match(rays, g["rays_interp"].view(clb.RAY_DTYPE), ("alpha", "U", "phi"))
    rays = g["rays_interp"].view(clb.RAY_DTYPE).copy()
    clb.rayprop_sphere(45.0, 15.0, 0.0, rays)
    assert_rays_match(rays, g["rays_prop1"].view(clb.RAY_DTYPE))
    rz = g["rz0"].view(clb.RAY_DTYPE).copy()          # alpha == 0 branch
    clb.rayprop_sphere(45.0, 15.0, 0.0, rz)
    assert_rays_match(rz, g["rz1"].view(clb.RAY_DTYPE))


def test_rays_multi_plane_vs_oracle(clb, oracle):
    rng = np.random.default_rng(21)
    order = 6
    npix = 12 << (2 * order)
    ro = oracle.init_rays(7, 15.0)[::5].copy()
    rg = ro.copy()
    w = [15.0 + 30.0 * p for p in range(6)]
    for p in range(4):
        maps = (rng.normal(size=(6, npix)) * np.array([1, 2e-4, 2e-4, 3e-3, 3e-3, 3e-3])[:, None]).astype(np.float32)
        for r in (ro, rg):
            r["alpha"] = 0; r["U"] = 0; r["phi"] = 0
        oracle.shearinterp(order, 3, maps, ro)
        clb.shearinterp_rays(maps, order, rg)
        wm1 = 0.0 if p == 0 else w[p - 1]
        oracle.rayprop(ro, w[p + 1], w[p], wm1)
        clb.rayprop_sphere(w[p + 1], w[p], wm1, rg)
        assert_rays_match(rg, ro)
    kap_o = 1 - 0.5 * (ro["A"][:, 0] + ro["A"][:, 3]); kap_g = 1 - 0.5 * (rg["A"][:, 0] + rg["A"][:, 3])
    assert np.abs(kap_g - kap_o).max() <= RAY_TOL * np.abs(kap_o).max()


def test_ray_init_matches_reference_init(clb, oracle):
    import torch
    from calclens_b200 import _lib
    from calclens_b200.rays import rays_from_device
    L = _lib.load()
    order = 6
    n = 12 << (2 * order)
    t = torch.empty(n * 176, dtype=torch.uint8, device="cuda")
    L.clb_ray_init_dev(t.data_ptr(), n, 0, order, 15.0, None)
    torch.cuda.synchronize()
    rg = rays_from_device(t)
    ro = oracle.init_rays(order, 15.0)
    assert np.array_equal(rg["nest"], ro["nest"])
    for f in ("beta", "n"):
        assert np.abs(rg[f] - ro[f]).max() <= 1e-15 * max(1.0, np.abs(ro[f]).max())
    assert np.array_equal(rg["A"], ro["A"]) and np.array_equal(rg["Aprev"], ro["Aprev"])


# ----------------------------------------------------------------------------------------------------------------
# whole lens plane through the C ABI, and the sharded (N > 1) decomposition emulated on one GPU
# ----------------------------------------------------------------------------------------------------------------
def _oracle_plane(oracle, order, lmax, counts, premul, densmul, backdens, rays, wpp1, wp, wpm1, w=None):
    m = ((counts * premul) * densmul - backdens).astype(np.float32)
    are, aim = oracle.map2alm(order, lmax, m, w)
    are, aim = oracle.poisson_filter(lmax, are, aim)
    maps = oracle.alm2allmaps(order, lmax, are, aim)
    rays["alpha"] = 0; rays["U"] = 0; rays["phi"] = 0
    oracle.shearinterp(order, 2, maps, rays)
    oracle.rayprop(rays, wpp1, wp, wpm1)
    return maps


def test_lens_plane_end_to_end(clb, oracle):
    from calclens_b200 import _lib
    L = _lib.load()
    order, lmax = 6, 128
    npix = 12 << (2 * order)
    rng = np.random.default_rng(31)
    counts = (8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32)
    premul, densmul, backdens = np.float32(1.0), np.float32(2e-4), np.float32(8.0 * np.exp(0.125) * 2e-4)
    ro = oracle.init_rays(order, 15.0); rg = ro.copy()
    plan = clb.HEALPixSHTPlan(order, lmax)
    for (wpp1, wp, wpm1) in ((45.0, 15.0, 0.0), (75.0, 45.0, 15.0)):
        _oracle_plane(oracle, order, lmax, counts, premul, densmul, backdens, ro, wpp1, wp, wpm1)
        L.clb_lens_plane(plan._h, counts.ctypes.data, float(premul), float(densmul), float(backdens), rg.ctypes.data, rg.size, wpp1, wp, wpm1)
        assert_rays_match(rg, ro)
    plan.destroy()


@pytest.mark.parametrize("nranks", [2, 3])
def test_sharded_decomposition_emulated_on_one_gpu(clb, oracle, nranks):
    """N ranks' plans on one device; the two all-to-alls are done by slicing the send buffers with the plans'
    per-peer counts.  The assembled result must equal the single-rank result bit for bit."""
    import torch
    order, lmax = 5, 64
    npix = 12 << (2 * order)
    rng = np.random.default_rng(41)
    m = rng.normal(size=npix).astype(np.float32)
    dm = torch.from_numpy(m).cuda()
    single = clb.HEALPixSHTPlan(order, lmax)
    s_re, s_im = single.legendre_analysis(single.ring_analysis(dm), poisson_filter=True)
    s_maps = single.ring_synthesis(single.legendre_synthesis(s_re, s_im)).cpu().numpy()
    s_re = s_re.cpu().numpy(); s_im = s_im.cpu().numpy()
    plans = [clb.HEALPixSHTPlan(order, lmax, nranks=nranks, rank=r) for r in range(nranks)]
    from calclens_b200 import layout
    rp_owner, m_owner = clb.default_owners(order, lmax, nranks)
    for r, p in enumerate(plans):    # the CUDA-free layout module and the C plan agree
        lay = layout.ExchangeLayout(1 << order, lmax, nranks, r, rp_owner, m_owner)
        assert p.counts == [lay.g_send_counts, lay.g_recv_counts, lay.b_send_counts, lay.b_recv_counts]
        assert list(p.m_local) == list(lay.my_m) and list(p.rp_local) == list(lay.my_rp)

    def all_to_all(sends, which_send, which_recv):
        recvs = []
        for d in range(nranks):
            parts = []
            for s in range(nranks):
                cnt = plans[s].counts[which_send]
                off = 2 * sum(cnt[:d])
                parts.append(sends[s][off:off + 2 * cnt[d]])
            recvs.append(torch.cat(parts) if parts else torch.zeros(0, dtype=torch.float64, device="cuda"))
            assert recvs[-1].numel() == 2 * sum(plans[d].counts[which_recv])
        return recvs
    g_send = [p.ring_analysis(dm).clone() for p in plans]
    g_recv = all_to_all(g_send, 0, 1)
    alms = [p.legendre_analysis(g, poisson_filter=True) for p, g in zip(plans, g_recv)]
    # reassemble alm in single-rank m-major order
    full_re = np.zeros_like(s_re); full_im = np.zeros_like(s_im)
    for p, (are, aim) in zip(plans, alms):
        are = are.cpu().numpy(); aim = aim.cpu().numpy(); off = 0
        for mm in p.m_local:
            n = lmax - mm + 1
            dst = clb.lm2index(mm, mm, lmax)
            full_re[dst:dst + n] = are[off:off + n]; full_im[dst:dst + n] = aim[off:off + n]; off += n
    assert np.array_equal(full_re, s_re) and np.array_equal(full_im, s_im)
    b_send = [p.legendre_synthesis(a[0], a[1]).clone() for p, a in zip(plans, alms)]
    b_recv = all_to_all(b_send, 2, 3)
    total = torch.zeros((6, npix), dtype=torch.float32, device="cuda")
    for p, b in zip(plans, b_recv):
        total += p.ring_synthesis(b)           # disjoint ring sets
    assert np.array_equal(total.cpu().numpy(), s_maps)
    for p in plans + [single]:
        p.destroy()


def test_multi_gpu_step_matches_single_gpu(clb):
    """two ranks over NCCL (needs >= 2 visible GPUs; skipped on a one-GPU box)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "dist_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("nranks", [2, 3])
def test_fused_exchange_emulated_on_one_gpu(clb, nranks):
    """The fused exchange (clb_sht_plan_set_peers): N ranks' plans on one device, every rank's producing kernels store
    straight into the owners' receive buffers (here plain device buffers standing in for NVLink peer mappings), maps
    assembled with clb_maps_broadcast_dev.  Result must equal the single-rank result bit for bit."""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    order, lmax = 5, 64
    npix = 12 << (2 * order)
    rng = np.random.default_rng(43)
    dm = torch.from_numpy(rng.normal(size=npix).astype(np.float32)).cuda()
    single = clb.HEALPixSHTPlan(order, lmax)
    s_re, s_im = single.legendre_analysis(single.ring_analysis(dm), poisson_filter=True)
    s_maps = single.ring_synthesis(single.legendre_synthesis(s_re, s_im)).cpu().numpy()
    plans = [clb.HEALPixSHTPlan(order, lmax, nranks=nranks, rank=r) for r in range(nranks)]
    g_send = [torch.full((2 * max(p.g_send_total, 1),), float("nan"), dtype=torch.float64, device="cuda") for p in plans]
    b_recv = [torch.full((2 * max(p.b_recv_total, 1),), float("nan"), dtype=torch.float64, device="cuda") for p in plans]
    maps = [torch.zeros((6, npix), dtype=torch.float32, device="cuda") for _ in plans]
    g_arr = (C.c_void_p * nranks)(*[t.data_ptr() for t in g_send])
    b_arr = (C.c_void_p * nranks)(*[t.data_ptr() for t in b_recv])
    peer_maps = (C.c_void_p * (6 * nranks))(*[maps[q][k].data_ptr() for q in range(nranks) for k in range(6)])
    for p in plans:
        L.clb_sht_plan_set_peers(p._h, g_arr, b_arr)
    for p, g in zip(plans, g_send):                   # every rank's ring FFT fills its own send buffer ...
        p.ring_analysis(dm, g)
    alms = []
    for p in plans:                                   # ... and every rank's Legendre analysis pulls from the owners' buffers
        are = torch.empty(max(p.Nlm, 1), dtype=torch.float64, device="cuda"); aim = torch.empty_like(are)
        L.clb_legendre_analysis_dev(p._h, None, are.data_ptr(), aim.data_ptr(), 1, None)
        alms.append((are, aim))
    for p, (are, aim) in zip(plans, alms):            # producers of b: every rank's Legendre synthesis
        L.clb_legendre_synthesis_dev(p._h, are.data_ptr(), aim.data_ptr(), None, None)
    for r, p in enumerate(plans):
        p.ring_synthesis(b_recv[r], maps[r])
    for r, p in enumerate(plans):
        ptrs = (C.c_void_p * 6)(*[maps[r][k].data_ptr() for k in range(6)])
        L.clb_maps_broadcast_dev(p._h, ptrs, peer_maps, None, 0, None)
    torch.cuda.synchronize()
    for r in range(nranks):
        assert not torch.isnan(g_send[r][:2 * plans[r].g_send_total]).any() and not torch.isnan(b_recv[r][:2 * plans[r].b_recv_total]).any()
        assert np.array_equal(maps[r].cpu().numpy(), s_maps), "rank %d maps differ from the single-rank maps" % r
    for p in plans + [single]:
        p.destroy()


def test_load_density_matches_scale_density(clb):
    """clb_load_density_dev (pinned host or device source, own rings) == copy + clb_scale_density_dev"""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    order = 6
    npix = 12 << (2 * order)
    rng = np.random.default_rng(5)
    host = torch.from_numpy((8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32)).pin_memory()
    pm, dmul, bd = float(np.float32(0.37)), float(np.float32(2.5e-4)), float(np.float32(1.1e-3))
    ref = host.cuda()
    L.clb_scale_density_dev(ref.data_ptr(), npix, pm, dmul, bd, None)
    plan = clb.HEALPixSHTPlan(order, 2 << order)
    for src in (host, host.cuda()):
        dst = torch.zeros(npix, dtype=torch.float32, device="cuda")
        L.clb_load_density_dev(plan._h, src.data_ptr(), dst.data_ptr(), pm, dmul, bd, None)
        torch.cuda.synchronize()
        assert torch.equal(dst, ref)
    # a rank of a sharded plan touches only its own rings
    p1 = clb.HEALPixSHTPlan(order, 2 << order, nranks=2, rank=1)
    dst = torch.full((npix,), -7.0, dtype=torch.float32, device="cuda")
    L.clb_load_density_dev(p1._h, host.data_ptr(), dst.data_ptr(), pm, dmul, bd, None)
    torch.cuda.synchronize()
    touched = dst != -7.0
    assert torch.equal(dst[touched], ref[touched]) and 0.3 < float(touched.float().mean()) < 0.7
    plan.destroy(); p1.destroy()


def test_solver_step_prefetch_and_fused_summary(clb):
    """clb_solver_step with the next plane prefetched from pinned host memory (clb_solver_set_next) gives the same rays as
    without, and the summary accumulated inside the ray kernel equals the separate summary kernel."""
    import torch
    from calclens_b200 import poisson
    order, lmax = 6, 128
    npix = 12 << (2 * order)
    rng = np.random.default_rng(17)
    maps = [torch.from_numpy((8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32)).pin_memory() for _ in range(3)]
    sc = (np.float32(1.0), np.float32(2e-4), np.float32(8.0 * np.exp(0.125) * 2e-4))
    planes = [(45.0, 15.0, 0.0), (75.0, 45.0, 15.0), (105.0, 75.0, 45.0)]
    a = poisson.LensPlaneSolver(order, lmax, order); a.init_rays(15.0)
    b = poisson.LensPlaneSolver(order, lmax, order); b.init_rays(15.0)
    for k, pl in enumerate(planes):
        sa = a.step(maps[k], *sc, *pl)
        nxt = (maps[k + 1],) + sc if k + 1 < len(planes) else None
        sb = b.step(maps[k], *sc, *pl, prefetch=nxt)
        # the six sums are accumulated with atomics (order varies) over values that largely cancel
        assert np.allclose(sa, sb, rtol=1e-9, atol=1e-13)
        ref = torch.zeros(6, dtype=torch.float64, device="cuda")
        a.lib.clb_ray_summary_dev(a.rays.data_ptr(), a.nrays, ref.data_ptr(), None)
        assert np.allclose(sa, ref.cpu().numpy(), rtol=1e-9, atol=1e-13)
    assert np.array_equal(a.rays_host().view(np.uint8), b.rays_host().view(np.uint8))
    a.close(); b.close()


def test_global_scratch_ring_fft_matches_shared_memory_path(clb):
    """Rings whose FFT work buffers exceed an SM's shared memory (r > 4095, i.e. Nside 8192) run from a global scratch
    buffer with persistent CTAs.  Forced on at a small Nside, that path must reproduce the shared-memory path bit for
    bit (same arithmetic, different memory)."""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    order, lmax = 6, 150
    npix = 12 << (2 * order)
    rng = np.random.default_rng(99)
    dm = torch.from_numpy(rng.normal(size=npix).astype(np.float32)).cuda()
    out = []
    try:
        for force in (0, 1):
            L.clb_set_tuning(4, force)
            plan = clb.HEALPixSHTPlan(order, lmax)
            g = plan.ring_analysis(dm)
            are, aim = plan.legendre_analysis(g, poisson_filter=True)
            b = plan.legendre_synthesis(are, aim)
            maps = plan.ring_synthesis(b)
            out.append((g.clone(), maps.clone()))
            plan.destroy()
    finally:
        L.clb_set_tuning(4, 0)
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


def test_ray_output_transform_vs_oracle(clb, oracle):
    """write_rays' pre-output transform (rayio.c:300-312) on the device vs the reference's paratrans_ray_curr2obs +
    rot_ray_ang2radec; the resident rays must stay untouched."""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    order = 5
    rng = np.random.default_rng(23)
    rays = oracle.init_rays(order, 15.0)
    rays["n"] += rng.normal(size=rays["n"].shape) * 0.02          # rays have drifted from their observed pixel
    for f, n in (("A", 4), ("Aprev", 4), ("U", 4), ("alpha", 2)):
        rays[f] += rng.normal(size=(rays.size, n)) * 0.05
    ro = rays.copy()
    oracle.ray_output(ro, order)
    d_in = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
    d_out = torch.zeros_like(d_in)
    L.clb_ray_output_dev(d_in.data_ptr(), d_out.data_ptr(), rays.size, order, None)
    torch.cuda.synchronize()
    rg = d_out.cpu().numpy().view(rays.dtype)
    assert np.array_equal(d_in.cpu().numpy(), rays.view(np.uint8).reshape(-1))
    assert np.array_equal(rg["nest"], ro["nest"]) and np.array_equal(rg["n"], ro["n"]) and np.array_equal(rg["U"], ro["U"])
    assert_rays_match(rg, ro)


def test_ngp_deposit_vs_oracle(clb, oracle):
    """NGP particle deposit (shtpoissonsolve.c:128-150): bit-exact pixel assignment; with equal-mass particles the
    float sums are order independent and must match the sequential reference loop bit for bit."""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    order = 6
    npix = 12 << (2 * order)
    rng = np.random.default_rng(29)
    n = 400000
    pos = rng.normal(size=(n, 3)).astype(np.float32) * np.float32(300.0)
    pos[:6] = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [1e-3, -1e-3, 1]]   # poles, axes, near-pole
    mass = np.full(n, 3.7e9, dtype=np.float32)
    want = oracle.deposit_ngp(pos, mass, order)
    dm = torch.zeros(npix, dtype=torch.float32, device="cuda")
    dp = torch.from_numpy(pos).cuda(); dmass = torch.from_numpy(mass).cuda()
    L.clb_deposit_ngp_dev(dp.data_ptr(), dmass.data_ptr(), n, order, dm.data_ptr(), None)
    torch.cuda.synchronize()
    got = dm.cpu().numpy()
    assert np.array_equal(got, want)
    assert abs(float(got.sum()) - n * 0.37) < 1e-3 * n * 0.37
    # unequal masses: same pixels, sums to float round-off
    mass2 = (3.7e9 * (0.5 + rng.random(n))).astype(np.float32)
    want2 = oracle.deposit_ngp(pos, mass2, order)
    dm.zero_()
    L.clb_deposit_ngp_dev(dp.data_ptr(), torch.from_numpy(mass2).cuda().data_ptr(), n, order, dm.data_ptr(), None)
    torch.cuda.synchronize()
    got2 = dm.cpu().numpy()
    assert np.array_equal(got2 != 0, want2 != 0)
    assert np.allclose(got2, want2, rtol=2e-6, atol=0)


def test_born_approximation_ray_step_vs_oracle(clb, oracle):
    """the -DBORNAPPRX form of rayprop_sphere (rayprop.c:40-62), mode bit 8"""
    order = 4
    rng = np.random.default_rng(37)
    rays = oracle.init_rays(order, 15.0)
    for f, n in (("A", 4), ("Aprev", 4), ("U", 4), ("alpha", 2)):
        rays[f] += rng.normal(size=(rays.size, n)) * 0.05
    ro = rays.copy(); rg = rays.copy()
    for (wp, wpm1, wpm2) in ((45.0, 15.0, 0.0), (75.0, 45.0, 15.0)):
        oracle.rayprop_born(ro, wp, wpm1, wpm2)
        clb.rayprop_sphere(wp, wpm1, wpm2, rg, born=True)
    assert_rays_match(rg, ro)
    assert np.array_equal(rg["beta"], ro["beta"]) and np.array_equal(rg["alpha"], ro["alpha"])


def test_point_mass_plane_baseline_config0(clb, oracle):
    """BASELINE configs[0]: the point-mass test plane (lensplanes/make_lensplanes_pointmass_test.c:156-192: one particle,
    NGP-deposited into one pixel) at Nside=256, lmax=512, full-sky rays at Nside=512 -- one lens plane through the C ABI
    against the reference, plus the analytic deflection |grad phi| = S/(4 pi) cot(d/2) (sign and
    scale; an unsmoothed pixel truncated at lmax rings, so the median over rays is compared, not every ray)."""
    from calclens_b200 import _lib
    L = _lib.load()
    order, lmax, ray_order = 8, 512, 9
    npix = 12 << (2 * order)
    counts = np.zeros(npix, dtype=np.float32)
    pix = 5 * npix // 12 + 1234                       # an equatorial-belt pixel away from the poles
    counts[pix] = 1.0
    premul, densmul, backdens = np.float32(1.0), np.float32(3e-2), np.float32(0.0)   # NOBACKDENS, as in the test build
    wpp1, wp, wpm1 = 45.0, 15.0, 0.0
    ro = oracle.init_rays(ray_order, wp); rg = ro.copy()
    ro["alpha"] = 0; ro["U"] = 0; ro["phi"] = 0
    m = ((counts * premul) * densmul - backdens).astype(np.float32)
    are, aim = oracle.map2alm(order, lmax, m)
    are, aim = oracle.poisson_filter(lmax, are, aim)
    maps = oracle.alm2allmaps(order, lmax, are, aim)
    oracle.shearinterp(order, 2, maps, ro)
    alpha_before_prop = ro["alpha"].copy()
    n_before = ro["n"].copy()
    oracle.rayprop(ro, wpp1, wp, wpm1)
    plan = clb.HEALPixSHTPlan(order, lmax)
    L.clb_lens_plane(plan._h, counts.ctypes.data, float(premul), float(densmul), float(backdens), rg.ctypes.data, rg.size,
                     wpp1, wp, wpm1)
    plan.destroy()
    assert_rays_match(rg, ro)
    # analytic check on the deflection the GPU path produced
    S = float(densmul) * (4.0 * np.pi / npix)         # integral of the scaled map over the sphere
    from oracle import port
    c = np.zeros(3)
    port.lib().port_nest2vec(port.ring2nest(pix, order), c.ctypes.data, order)
    nhat = n_before / np.linalg.norm(n_before, axis=1)[:, None]
    d = np.arccos(np.clip(nhat @ c, -1, 1))
    sel = (d > 0.1) & (d < 2.5)
    amp = np.linalg.norm(rg["alpha"], axis=1)         # alpha = -grad phi (shtpoissonsolve.c:693-694)
    ratio = amp[sel] / (S / (4 * np.pi) / np.tan(d[sel] / 2))
    assert 0.97 < np.median(ratio) < 1.03, np.median(ratio)
    assert np.allclose(rg["alpha"], alpha_before_prop, rtol=0, atol=1e-8 * np.abs(alpha_before_prop).max())


@pytest.mark.parametrize("order,lmax", [(5, 77), (6, 191), (3, 9)])
def test_stage_outputs_stay_inside_their_buffers(clb, order, lmax):
    """Every stage writes exactly its output buffer: canaries on both sides of g, alm, b and the maps must survive, and
    no output element may be left unwritten (ragged lmax: partial degree blocks and coefficient tiles)."""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    npix = 12 << (2 * order)
    plan = clb.HEALPixSHTPlan(order, lmax)
    pad = 4096
    rng = np.random.default_rng(7)

    def guarded(n, dtype, fill=float("nan")):
        t = torch.full((n + 2 * pad,), -12345.0, dtype=dtype, device="cuda")
        t[pad:pad + n] = fill
        return t, t[pad:pad + n]

    def check(t, n, what):
        assert bool((t[:pad] == -12345.0).all()) and bool((t[pad + n:] == -12345.0).all()), "%s: write outside the buffer" % what
        assert not bool(torch.isnan(t[pad:pad + n]).any()), "%s: element left unwritten" % what

    dm = torch.from_numpy(rng.normal(size=npix).astype(np.float32)).cuda()
    g_all, g = guarded(2 * plan.g_send_total, torch.float64)
    are_all, are = guarded(plan.Nlm, torch.float64)
    aim_all, aim = guarded(plan.Nlm, torch.float64)
    b_all, b = guarded(2 * plan.b_send_total, torch.float64)
    maps_all, maps = guarded(6 * npix, torch.float32)
    plan.ring_analysis(dm, g)
    plan.legendre_analysis(g, are, aim, poisson_filter=True)
    plan.legendre_synthesis(are, aim, b)
    plan.ring_synthesis(b, maps.view(6, npix))
    torch.cuda.synchronize()
    check(g_all, 2 * plan.g_send_total, "g"); check(are_all, plan.Nlm, "alm_re"); check(aim_all, plan.Nlm, "alm_im")
    check(b_all, 2 * plan.b_send_total, "b"); check(maps_all, 6 * npix, "maps")
    plan.destroy()


def test_next_rows_golden(clb):
    """the three 'next' rows against the committed golden vectors (no oracle library needed)"""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    g = np.load(os.path.join(GOLD, "next_rows.npz"))
    order = int(g["dep_order"])
    dm = torch.zeros(12 << (2 * order), dtype=torch.float32, device="cuda")
    dp = torch.from_numpy(g["dep_pos"]).cuda(); dmass = torch.from_numpy(g["dep_mass"]).cuda()
    L.clb_deposit_ngp_dev(dp.data_ptr(), dmass.data_ptr(), int(dmass.numel()), order, dm.data_ptr(), None)
    torch.cuda.synchronize()
    assert np.array_equal(dm.cpu().numpy(), g["dep_map"])
    rays = g["rays_in"].copy().view(clb.RAY_DTYPE)
    d_in = torch.from_numpy(rays.view(np.uint8).copy()).cuda(); d_out = torch.zeros_like(d_in)
    L.clb_ray_output_dev(d_in.data_ptr(), d_out.data_ptr(), rays.size, 3, None)
    torch.cuda.synchronize()
    assert_rays_match(d_out.cpu().numpy().view(clb.RAY_DTYPE), g["rays_output"].view(clb.RAY_DTYPE))
    rb = rays.copy()
    clb.rayprop_sphere(45.0, 15.0, 0.0, rb, born=True); clb.rayprop_sphere(75.0, 45.0, 15.0, rb, born=True)
    assert_rays_match(rb, g["rays_born"].view(clb.RAY_DTYPE))


# ----------------------------------------------------------------------------------------------------------------
# round 2: the kernel's own stencil path, halo masks, the C solver with emulated ranks, parity at the quoted sizes
# ----------------------------------------------------------------------------------------------------------------
def _edge_vectors(order, rng, n_random):
    """unit vectors: random ones plus points within a few ulp of ring boundaries (theta of every sampled ring), of pixel
    boundaries in phi, of the poles and of phi = 0 / 2 pi"""
    nside = 1 << order
    v = rng.normal(size=(n_random, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    rings = np.unique(np.concatenate([np.arange(1, min(4 * nside, 40)), rng.integers(1, 4 * nside, 300),
                                      [nside - 1, nside, nside + 1, 2 * nside, 3 * nside - 1, 3 * nside, 3 * nside + 1, 4 * nside - 1]]))
    rings = rings[(rings >= 1) & (rings <= 4 * nside - 1)]
    zs = []
    for r in rings:
        if r < nside:
            z = 1.0 - r * r / (3.0 * nside * nside)
        elif r <= 3 * nside:
            z = (2 * nside - r) * 2.0 / (3.0 * nside)
        else:
            rr = 4 * nside - r
            z = -(1.0 - rr * rr / (3.0 * nside * nside))
        zs.append(z)
    zs = np.array(zs)
    edge = []
    for k in (-2, -1, 0, 1, 2):
        z = zs.copy()
        for _ in range(abs(k)):
            z = np.nextafter(z, np.inf if k > 0 else -np.inf)
        z = np.clip(z, -1.0, 1.0)
        npx = np.where(rings < nside, 4 * rings, np.where(rings <= 3 * nside, 4 * nside, 4 * (4 * nside - rings)))
        j = rng.integers(0, npx)
        for frac in (0.0, 0.5):                                  # pixel centres / pixel edges of shifted and unshifted rings
            ph = (j + frac) * 2.0 * np.pi / npx
            for dk in (-1, 0, 1):
                p = ph.copy()
                if dk:
                    p = np.nextafter(p, np.inf if dk > 0 else -np.inf)
                s = np.sqrt(np.maximum(0.0, (1.0 - z) * (1.0 + z)))
                edge.append(np.stack([s * np.cos(p), s * np.sin(p), z], axis=1))
    edge = np.concatenate(edge)
    poles = np.array([[0, 0, 1.0], [0, 0, -1.0], [1e-300, 0, 1.0], [1e-9, 1e-9, 1.0], [-1e-9, 1e-12, -1.0], [1, 0, 0], [1, -1e-17, 0],
                      [1, 1e-17, 0], [-1, 1e-17, 0], [-1, -1e-17, 0], [0, 1, 0], [0, -1, 0]], dtype=np.float64)
    return np.concatenate([poles, edge, v])


@pytest.mark.parametrize("order", [3, 8, 12])
def test_ray_kernel_stencil_bit_exact_1e7(clb, oracle, order):
    """The stencil exactly as ray_step_kernel forms it (vec2ang + get_interpol_tab through clb_ray_stencil_dev) against
    (a) the device mirror of the reference's get_interpol on ~1e7 points (indices identical everywhere, weights equal to
    1e-9) and (b) the reference itself (oracle get_interpol, host libm) on 60000 of them: ZERO index mismatches on the
    random points.  The adversarial points sit within 0-2 ulp of a ring or pixel boundary, where one ulp of cos/atan2
    (device libm vs glibc) decides between two neighbouring stencils; there a different stencil is accepted
    only if it is the same interpolant, i.e. the pixels the two stencils do not share carry weight < 1e-9."""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(1000 + order)
    nrand = 10_000_000 if order == 12 else 2_000_000
    v = _edge_vectors(order, rng, nrand)
    n = v.shape[0]
    tv = torch.from_numpy(np.ascontiguousarray(v)).cuda()
    pa = torch.empty((n, 4), dtype=torch.int64, device="cuda"); wa = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    pb = torch.empty_like(pa); wb = torch.empty_like(wa)
    L.clb_ray_stencil_dev(order, n, tv.data_ptr(), pa.data_ptr(), wa.data_ptr(), None)
    L.clb_healpix_interpol_dev(order, n, tv.data_ptr(), pb.data_ptr(), wb.data_ptr(), None)
    torch.cuda.synchronize()
    assert torch.equal(pa, pb), "table path and reference-mirror path disagree on %d stencils" % int((pa != pb).any(dim=1).sum())
    assert float((wa - wb).abs().max()) < 1e-9
    # against the reference's own get_interpol (host): all edge points + a random sample
    nedge = n - nrand
    idx = np.concatenate([np.arange(nedge), nedge + rng.integers(0, nrand, 60000 - min(nedge, 30000))])[:60000]
    pix = pa[torch.from_numpy(idx).cuda()].cpu().numpy(); wgt = wa[torch.from_numpy(idx).cuda()].cpu().numpy()
    import ctypes
    Lr = getattr(oracle, "lib")()
    has_ref = hasattr(Lr, "vec2ang")
    bad_random, edge_flips = 0, 0
    for k, i in enumerate(idx):
        if has_ref:
            th, ph = ctypes.c_double(), ctypes.c_double()
            vv = (ctypes.c_double * 3)(*v[i])
            Lr.vec2ang(vv, ctypes.byref(th), ctypes.byref(ph))
            theta, phi = th.value, ph.value
        else:
            theta = np.arctan2(np.sqrt(v[i, 0] ** 2 + v[i, 1] ** 2), v[i, 2]); phi = np.arctan2(v[i, 1], v[i, 0]) % (2 * np.pi)
        p, w = oracle.get_interpol(theta, phi, order)
        if list(pix[k]) == p:
            assert np.abs(wgt[k] - np.array(w)).max() < 1e-9
            continue
        if i >= nedge:
            bad_random += 1
            continue
        d = {}
        for q, ww in zip(pix[k], wgt[k]):
            d[int(q)] = d.get(int(q), 0.0) + float(ww)
        for q, ww in zip(p, w):
            d[int(q)] = d.get(int(q), 0.0) - float(ww)
        assert sum(abs(x) for x in d.values()) < 1e-9, "edge point %d: the two stencils are different interpolants %r" % (i, d)
        edge_flips += 1
    assert bad_random == 0, "%d random stencils differ from the reference's get_interpol" % bad_random
    print("order %d: %d points on the device, %d checked against the reference; %d of %d boundary points picked the neighbouring "
          "(equivalent) stencil" % (order, n, len(idx), edge_flips, min(nedge, len(idx))))


def _solver_inputs(order, seed):
    import torch
    npix = 12 << (2 * order)
    rng = np.random.default_rng(seed)
    counts = torch.from_numpy((8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32)).pin_memory()
    sc = (np.float32(1.0), np.float32(2e-4), np.float32(8.0 * np.exp(0.125) * 2e-4))
    return counts, sc


@pytest.mark.parametrize("nranks,order,halo", [(2, 5, 1.0), (3, 6, 0.5), (4, 6, 0.0)])
def test_c_solver_emulated_ranks_with_halo_masks(clb, nranks, order, halo):
    """clb_solver_* with N ranks emulated as threads on one GPU: the fused exchange, the halo-limited map broadcast with
    REAL need masks (order == coarse order 5 is the corner-cut case of the 4-pixel groups) and the per-stencil mask
    check of the ray kernel.  Rays and summaries must equal the single-rank solver bit for bit; every pixel any ray's
    stencil touches must have been delivered; pushing a ray out of its domain must raise the error flag."""
    import torch
    from calclens_b200 import poisson
    from tests.emu import ThreadRanks
    lmax = 2 << order
    counts, sc = _solver_inputs(order, 77)
    planes = [(45.0, 15.0, 0.0), (75.0, 45.0, 15.0)]
    single = poisson.LensPlaneSolver(order, lmax, order); single.init_rays(15.0)
    centres = single.rays_host()["n"].copy()      # pixel centres (x 15) of every NEST pixel at this order
    ssum = [single.step(counts, *sc, *pl) for pl in planes]
    srays = single.rays_host().copy()
    smaps = single.maps.cpu().numpy()
    emu = ThreadRanks(nranks)

    def body(rank, gather):
        s = poisson.LensPlaneSolver(order, lmax, order, nranks=nranks, rank=rank, allgather=gather, halo_deg=halo)
        assert s.fused
        assert (s._need is not None) == (halo > 0)
        s.init_rays(15.0)
        sums = [s.step(counts, *sc, *pl) for pl in planes]
        rays = s.rays_host().copy()
        maps = s.maps.cpu().numpy()
        need = None if s._need is None else s._need.cpu().numpy()
        # move one ray to the centre of a coarse cell this rank does NOT receive: its stencil is outside domain + halo
        err = 0
        if need is not None and nranks > 1:
            far = np.nonzero(((need >> rank) & 1) == 0)[0]
            assert far.size > 0, "rank %d receives the whole sky: the mask test is vacuous" % rank
            pix_far = int(far[far.size // 2]) << (2 * (order - 5))
            rr = rays[:1].copy(); rr["n"] = centres[pix_far]
            s.rays[:176] = torch.from_numpy(rr.view(np.uint8)).cuda()
            s.lib.clb_solver_ray_update(s._cs, 105.0, 75.0, 45.0, 1 | 2, 0, None)
            err = s.lib.clb_solver_check(s._cs, None)
        first = s.first_nest
        s.close()
        return sums, rays, maps, need, err, first

    res = emu.run(body)
    got = np.concatenate([r[1] for r in res])
    assert np.array_equal(got.view(np.uint8), srays.view(np.uint8)), "rays differ from the single-rank solver"
    for k in range(len(planes)):
        tot = sum(r[0][k] for r in res)
        assert np.allclose(tot, ssum[k], rtol=1e-9, atol=1e-13)
    if halo > 0:
        co = 5
        npix = 12 << (2 * order)
        pix = torch.arange(npix, dtype=torch.int64, device="cuda"); out = torch.empty_like(pix)
        single.lib.clb_healpix_index_dev(0, order, npix, pix.data_ptr(), None, None, out.data_ptr(), None)   # ring -> nest
        torch.cuda.synchronize()
        cell = out.cpu().numpy() >> (2 * (order - co))
        for q, (_, rays, maps, need, err, first) in enumerate(res):
            assert err & 1, "rank %d: a ray far outside domain + halo did not raise the error flag" % q
            # every map pixel inside the cells this rank needs equals the single-rank map (nothing undelivered)
            needed = ((need[cell] >> q) & 1) == 1
            assert needed.any()
            assert np.array_equal(maps[:, needed].view(np.uint32), smaps[:, needed].view(np.uint32)), "rank %d misses needed pixels" % q
    else:
        for q, r in enumerate(res):
            assert np.array_equal(r[2].view(np.uint32), smaps.view(np.uint32)), "rank %d: full maps differ" % q
    single.close()


def test_tuning_fallback_small_maps(clb, oracle):
    """clb_set_tuning(1, 6) at Nside 64 (96 <= ring pairs < 192): the rings-per-thread fallback must land on an
    instantiated kernel and analyse every ring pair."""
    from calclens_b200 import _lib
    L = _lib.load()
    order, lmax = 6, 128
    rng = np.random.default_rng(8)
    m = rng.normal(size=12 << (2 * order)).astype(np.float32)
    are, aim = oracle.map2alm(order, lmax, m)
    try:
        for R in (6, 4, 2, 1):
            L.clb_set_tuning(1, R)
            plan = clb.HEALPixSHTPlan(order, lmax)
            gre, gim = clb.map2alm_mpi(m, plan)
            assert alm_err(gre, gim, are, aim) <= ALM_TOL, R
            plan.destroy()
    finally:
        L.clb_set_tuning(1, 8)
