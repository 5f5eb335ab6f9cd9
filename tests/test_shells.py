"""GPU tests of the two-shells-per-pass Legendre kernels and of the solver's pair mode (the
reference solves plane by plane, shtpoissonsolve.c:517-570, so the bar is: every shell of a batched pass equals the
one-shell result BIT FOR BIT, and the one-shell result is the one the oracle tests pin).  Also the partial-sum row
layout of the Legendre analysis (a warp walks several ring chunks and owns one row): any row count must give the oracle's
alm."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALM_TOL = 1e-10


@pytest.fixture(scope="module")
def clb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import calclens_b200
    from calclens_b200 import _lib
    assert _lib.load().clb_device_count() >= 1
    return calclens_b200


def _two_g(plan, seed):
    import torch
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    g = torch.empty(2 * 2 * max(plan.g_send_total, 1), dtype=torch.float64, device="cuda")
    maps = []
    for s in range(2):
        m = torch.randn(plan.npix, device="cuda", dtype=torch.float32, generator=gen) * (1.0 + s)
        maps.append(m)
        plan.ring_analysis(m, g[2 * s * plan.g_send_total:])
    return maps, g


@pytest.mark.parametrize("order,lmax", [(1, 5), (2, 8), (4, 47), (5, 64), (6, 191), (7, 256), (8, 512), (9, 1024)])
def test_two_shell_legendre_equals_one_shell_bit_for_bit(clb, order, lmax):
    import torch
    plan = clb.HEALPixSHTPlan(order, lmax)
    _, g = _two_g(plan, 100 + order)
    n, gt, bt = plan.Nlm, plan.g_send_total, plan.b_send_total
    are2, aim2 = plan.legendre_analysis(g, poisson_filter=True, nshell=2)
    b2 = plan.legendre_synthesis(are2, aim2, nshell=2)
    for s in range(2):
        are1, aim1 = plan.legendre_analysis(g[2 * s * gt:], poisson_filter=True)
        assert torch.equal(are1[:n], are2[s * n:(s + 1) * n]) and torch.equal(aim1[:n], aim2[s * n:(s + 1) * n]), "alm of shell %d" % s
        b1 = plan.legendre_synthesis(are1, aim1)
        assert torch.equal(b1[:2 * bt], b2[2 * s * bt:2 * (s + 1) * bt]), "b of shell %d" % s
    assert float(are2[:n].abs().max()) > 0 and not torch.equal(are2[:n], are2[n:2 * n])
    plan.destroy()


@pytest.mark.parametrize("order,lmax", [(6, 128), (8, 512), (10, 2048)])
def test_two_shell_legendre_other_rings_per_thread(clb, order, lmax):
    """the tunable rings-per-thread variants of the two-shell kernels: same alm to FP64 round-off (the ring chunks are cut
    differently, so the sums associate differently), b of a given alm bit for bit"""
    import torch
    from calclens_b200 import _lib
    L = _lib.load()
    plan = clb.HEALPixSHTPlan(order, lmax)
    _, g = _two_g(plan, 7)
    n = plan.Nlm
    ref = plan.legendre_analysis(g, poisson_filter=True, nshell=2)
    ref = (ref[0].clone(), ref[1].clone())
    bref = plan.legendre_synthesis(ref[0], ref[1], nshell=2).clone()
    scale = float(torch.sqrt((ref[0] ** 2 + ref[1] ** 2).sum()))
    try:
        for R in (6, 4, 2, 1):
            L.clb_set_tuning(10, R)
            a = plan.legendre_analysis(g, poisson_filter=True, nshell=2)
            err = float(torch.sqrt(((a[0] - ref[0]) ** 2 + (a[1] - ref[1]) ** 2).sum())) / scale
            assert err < 1e-13, (R, err)
        for R in (4, 2, 1):
            L.clb_set_tuning(9, R)
            b = plan.legendre_synthesis(ref[0], ref[1], nshell=2)
            assert torch.equal(b, bref), R
    finally:
        L.clb_set_tuning(10, 8); L.clb_set_tuning(9, 3)
    plan.destroy()


@pytest.mark.parametrize("order,lmax", [(5, 64), (8, 512)])
def test_analysis_partial_row_counts_vs_oracle(clb, oracle, order, lmax):
    """rows = 1 (one warp walks every chunk of an m), 2, 3 and automatic: all within the alm tolerance of the oracle"""
    from calclens_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(order)
    m = rng.normal(size=12 << (2 * order)).astype(np.float32)
    are, aim = oracle.map2alm(order, lmax, m)
    scale = np.sqrt((are ** 2 + aim ** 2).sum())
    try:
        for rows in (1, 2, 3, 0):
            L.clb_set_tuning(5, rows)
            plan = clb.HEALPixSHTPlan(order, lmax)
            gre, gim = clb.map2alm_mpi(m, plan)
            err = np.sqrt(((gre - are) ** 2 + (gim - aim) ** 2).sum()) / scale
            assert err <= ALM_TOL, (rows, err)
            plan.destroy()
    finally:
        L.clb_set_tuning(5, 0)


def _planes(order, nplanes, seed):
    import torch
    npix = 12 << (2 * order)
    rng = np.random.default_rng(seed)
    maps = [torch.from_numpy((8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32)).pin_memory() for _ in range(nplanes)]
    sc = [(np.float32(1.0 + 0.1 * p), np.float32(2e-4), np.float32(8.0 * np.exp(0.125) * 2e-4 * (1.0 + 0.1 * p))) for p in range(nplanes)]
    w = [(30.0 * p + 45.0, 30.0 * p + 15.0, 0.0 if p == 0 else 30.0 * p - 15.0) for p in range(nplanes)]
    return maps, sc, w


@pytest.mark.parametrize("order,where", [(5, "host"), (6, "device"), (7, "host")])
def test_solver_pair_mode_equals_plane_by_plane(clb, order, where):
    """five planes: plane by plane, and as pairs (0,1) (2,3) + a single -- with the partner's density prefetched through
    clb_solver_set_next or loaded on the spot.  Rays byte for byte, the six sums of every plane, and the maps of both sets."""
    import torch
    from calclens_b200 import poisson
    lmax = 2 << order
    maps, sc, w = _planes(order, 5, 3)
    if where == "device":
        maps = [m.cuda() for m in maps]
    a = poisson.LensPlaneSolver(order, lmax, order); a.init_rays(15.0)
    sums_a = [a.step(maps[p], *sc[p], *w[p]) for p in range(5)]
    rays_a = a.rays_host().copy()
    maps4 = a.maps.cpu().numpy().copy()
    b = poisson.LensPlaneSolver(order, lmax, order); b.init_rays(15.0)
    assert b.shells == 2
    sums_b = []
    for p in range(5):
        pair = (maps[p + 1],) + tuple(sc[p + 1]) if p in (0, 2) else None
        # prefetch what is to come: once the ring FFTs of a pair step are enqueued both density buffers are free again
        pre = [(maps[q],) + tuple(sc[q]) for q in range(p + 2, min(p + 4, 5))] if (order != 7 and pair) else None
        launches0 = b.lib.clb_solver_query(b._cs, 3)
        sums_b.append(b.step(maps[p], *sc[p], *w[p], pair=pair, prefetch=pre or None))
        if p in (1, 3):
            # the rays-only step launches the ray kernel (+ at most two density prefetches), no SHT stage
            assert b.lib.clb_solver_query(b._cs, 3) - launches0 <= 3
            if p == 3:
                m3 = b.maps2.cpu().numpy().copy()
    rays_b = b.rays_host().copy()
    assert np.array_equal(rays_a.view(np.uint8), rays_b.view(np.uint8)), "rays differ between pair mode and plane-by-plane"
    for p in range(5):
        assert np.allclose(sums_a[p], sums_b[p], rtol=1e-9, atol=1e-13), p   # (atomic accumulation order varies)
    assert np.array_equal(b.maps.cpu().numpy().view(np.uint32), maps4.view(np.uint32))   # plane 4 was solved alone, map set 0
    # plane 3 was the partner of the pass (2, 3): its maps sit in the second set and equal a one-shell solve of plane 3
    a.load_density(maps[3], *sc[3]); a.solve(); torch.cuda.synchronize()
    assert np.array_equal(a.maps.cpu().numpy().view(np.uint32), m3.view(np.uint32)), "second map set differs from the one-shell solve"
    a.close(); b.close()


@pytest.mark.parametrize("nranks,order,halo", [(2, 5, 1.0), (3, 6, 0.0)])
def test_solver_pair_mode_emulated_ranks(clb, nranks, order, halo):
    """the fused exchange with two shells per pass (second halves of the peer buffers, second map set broadcast), ranks
    emulated as threads on one GPU: rays equal the single-rank plane-by-plane solver bit for bit"""
    from calclens_b200 import poisson
    from tests.emu import ThreadRanks
    lmax = 2 << order
    maps, sc, w = _planes(order, 4, 11)
    single = poisson.LensPlaneSolver(order, lmax, order); single.init_rays(15.0)
    ssum = [single.step(maps[p], *sc[p], *w[p]) for p in range(4)]
    srays = single.rays_host().copy()
    single.close()
    emu = ThreadRanks(nranks)

    def body(rank, gather):
        s = poisson.LensPlaneSolver(order, lmax, order, nranks=nranks, rank=rank, allgather=gather, halo_deg=halo)
        assert s.fused and s.shells == 2
        s.init_rays(15.0)
        sums = []
        for p in range(4):
            pair = (maps[p + 1],) + tuple(sc[p + 1]) if p in (0, 2) else None
            sums.append(s.step(maps[p], *sc[p], *w[p], pair=pair))
        rays = s.rays_host().copy()
        s.close()
        return sums, rays

    res = emu.run(body)
    got = np.concatenate([r[1] for r in res])
    assert np.array_equal(got.view(np.uint8), srays.view(np.uint8)), "rays differ from the single-rank solver"
    for k in range(4):
        tot = sum(r[0][k] for r in res)
        assert np.allclose(tot, ssum[k], rtol=1e-9, atol=1e-13)
