"""CPU-side host logic: indexing helpers, ring-weight reader, plane geometry, exchange layouts, work model."""
import os

import numpy as np

import calclens_b200 as clb
from calclens_b200 import layout, poisson

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_lm_indexing():
    lmax = 9
    assert clb.num_lms(lmax) == (lmax + 1) * (lmax + 2) // 2
    k = 0
    for m in range(lmax + 1):
        for l in range(m, lmax + 1):
            assert clb.lm2index(l, m, lmax) == k
            k += 1
    assert clb.order2lmax(4) == 47


def test_ring_weight_reader_matches_reference_reader():
    ref_dir = "/root/reference/healpix_weights"
    if not os.path.isdir(ref_dir):
        import pytest
        pytest.skip("reference data files not present")
    g = np.load(os.path.join(GOLD, "ring_weights.npz"))
    for order in (1, 4, 8):
        assert np.array_equal(clb.read_ring_weights(ref_dir, order), g["n%05d" % (1 << order)])


def test_plane_params_follow_set_plane_params():
    c = poisson.Cosmology(0.27)
    p0 = poisson.plane_params(0, 50, 1500.0, 0.27, c)
    assert p0["wpm1"] == 0.0 and p0["wp"] == 15.0 and p0["wpp1"] == 45.0
    pl = poisson.plane_params(49, 50, 1500.0, 0.27, c)
    assert pl["wpp1"] == 1500.0 and abs(pl["wp"] - 1485.0) < 1e-12
    # densfact * Omega_m rho_crit V_rad == backdens (the mean density maps to the background term)
    for p in (p0, pl):
        binL, w = p["binL"], p["wp"]
        vrad = ((w + binL / 2) ** 3 - (w - binL / 2) ** 3) / 3.0
        assert abs(p["densfact"] * 0.27 * poisson.RHO_CRIT * vrad / p["backdens"] - 1.0) < 1e-12
    # comoving distance table: w(a=1) = 0, monotone, EdS limit check at Om = 1: w = 2*2997.9*(1-sqrt(a))
    e = poisson.Cosmology(1.0)
    a = 0.25
    i = int((a - e.AMIN) / (e.AMAX - e.AMIN) * (e.N - 1))
    assert abs(e.comv[i] - 2 * 2997.92458 * (1 - np.sqrt(e.aexpn[i]))) < 1e-6
    assert abs(e.acomvdist(2 * 2997.92458 * (1 - np.sqrt(0.25))) - 0.25) < 1e-6


def test_density_scalings_are_float32():
    s = poisson.density_scalings(8, 3.2e10, 1e-19, 2e-4)
    assert all(isinstance(x, np.float32) for x in s)
    assert s[0] == np.float32(3.2)


def test_default_owners_balanced():
    for order, lmax, nranks in ((6, 128, 2), (8, 512, 8), (10, 2048, 4)):
        rp, mo = clb.default_owners(order, lmax, nranks)
        assert rp.size == 2 << order and mo.size == lmax + 1
        cnt = np.bincount(rp, minlength=nranks)
        assert cnt.max() - cnt.min() <= 4
        # pixel balance (polar rings are shorter): within 10 %
        nside = 1 << order
        npx = np.minimum(np.arange(1, 2 * nside + 1), nside) * 4
        load = np.bincount(rp, weights=npx, minlength=nranks)
        assert load.max() / load.min() < 1.1
        assert np.bincount(mo, minlength=nranks).max() - np.bincount(mo, minlength=nranks).min() <= 1


def test_exchange_layout_is_a_consistent_transpose():
    """what rank s sends to rank d must be exactly what d expects from s, element by element"""
    order, lmax, nranks = 3, 12, 3
    nside = 1 << order
    rp_owner, m_owner = clb.default_owners(order, lmax, nranks)
    L = [layout.ExchangeLayout(nside, lmax, nranks, r, rp_owner, m_owner) for r in range(nranks)]
    for s in range(nranks):
        for d in range(nranks):
            assert L[s].g_send_counts[d] == L[d].g_recv_counts[s]
            assert L[s].b_send_counts[d] == L[d].b_recv_counts[s]
    # simulate the all-to-all with tagged values
    for kind in ("g", "b"):
        send = []
        for r in range(nranks):
            tot = sum(L[r].g_send_counts) if kind == "g" else sum(L[r].b_send_counts)
            buf = np.full(tot, -1, dtype=np.int64)
            if kind == "g":
                for m in range(lmax + 1):
                    for rp in L[r].my_rp:
                        for h in (0, 1):
                            buf[L[r].g_send_index(m, rp, h)] = (m * 1000 + rp) * 2 + h
            else:
                for m in L[r].my_m:
                    for f in range(6):
                        for rp in range(2 * nside):
                            for h in (0, 1):
                                buf[L[r].b_send_index(m, f, rp, h)] = ((m * 1000 + rp) * 2 + h) * 6 + f
            assert (buf >= 0).all()
            send.append(buf)
        for d in range(nranks):
            sb = [L[s].g_sbase if kind == "g" else L[s].b_sbase for s in range(nranks)]
            recv = np.concatenate([send[s][sb[s][d]:sb[s][d + 1]] for s in range(nranks)])
            if kind == "g":
                for m in L[d].my_m:
                    for rp in range(2 * nside):
                        for h in (0, 1):
                            assert recv[L[d].g_recv_index(m, rp, h)] == (m * 1000 + rp) * 2 + h
            else:
                for m in range(lmax + 1):
                    for f in range(6):
                        for rp in L[d].my_rp:
                            for h in (0, 1):
                                assert recv[L[d].b_recv_index(m, f, rp, h)] == ((m * 1000 + rp) * 2 + h) * 6 + f


def test_ray_ranges_partition():
    for nranks in (1, 2, 3, 8):
        r = layout.ray_ranges(4, nranks)
        assert r[0][0] == 0 and r[-1][1] == 12 << 8
        assert all(r[i][1] == r[i + 1][0] for i in range(nranks - 1))


def test_triple_count_matches_survey_table():
    import bench
    # Nside 256 / lmax 512 -> 5.60e7 triples with the cut
    t = bench.triple_count(256, 512)
    assert abs(t / 5.60e7 - 1) < 0.01


def test_domain_masks_cover_domains_and_halo():
    """clb_domain_masks (host code of the library, no GPU): every coarse cell is needed by the rank that owns its rays,
    margins only ever add ranks, a zero margin gives exactly the owners, and the needed fraction stays near 1/N + halo."""
    import ctypes as C
    import math
    from calclens_b200 import _lib
    L = _lib.load()
    ray_order, co = 9, 4
    nc = 12 << (2 * co)
    nray = 12 << (2 * ray_order)
    for nranks in (2, 3, 8):
        own = np.zeros(nc, dtype=np.uint8)
        L.clb_domain_masks(ray_order, nranks, co, 0.0, own.ctypes.data)
        lo = np.arange(nc, dtype=np.int64) << (2 * (ray_order - co))
        hi = lo + (1 << (2 * (ray_order - co)))
        for q in range(nranks):
            qlo, qhi = (nray * q) // nranks, (nray * (q + 1)) // nranks
            expect = (lo < qhi) & (qlo < hi)
            assert np.array_equal(((own >> q) & 1).astype(bool), expect)
        wide = np.zeros(nc, dtype=np.uint8)
        L.clb_domain_masks(ray_order, nranks, co, math.radians(8.0), wide.ctypes.data)
        assert np.all((wide & own) == own) and np.any(wide != own)
        frac = np.unpackbits(wide[:, None], axis=1).sum() / (nc * nranks)
        assert 1.0 / nranks < frac < 1.0 / nranks + 0.45


def test_fused_exchange_addresses_match_the_all_to_all_layout():
    """The peer-memory exchange reads/writes exactly the elements the all-to-all-v would have delivered: the pull address
    of g inside the ring owner's send buffer equals that owner's own g_send_index, and the push address of b inside the
    ring owner's receive buffer equals that owner's own b_recv_index (CUDA-free mirror of sht_plan_set_peers)."""
    from calclens_b200 import layout, sht
    order, lmax = 4, 37
    nside = 1 << order
    for nranks in (2, 3, 5):
        rp_owner, m_owner = sht.default_owners(order, lmax, nranks)
        lays = [layout.ExchangeLayout(nside, lmax, nranks, r, rp_owner, m_owner) for r in range(nranks)]
        rng = np.random.default_rng(nranks)
        for _ in range(400):
            m = int(rng.integers(0, lmax + 1)); rp = int(rng.integers(0, 2 * nside)); hemi = int(rng.integers(0, 2))
            f = int(rng.integers(0, 6))
            me = lays[int(m_owner[m])]                       # the rank whose Legendre stage handles m
            m_idx = int(me.m_local[m])
            q, off = me.g_pull_index(m_idx, rp, hemi)
            assert q == rp_owner[rp] and off == lays[q].g_send_index(m, rp, hemi)
            q, off = me.b_push_index(m_idx, f, rp, hemi)
            assert q == rp_owner[rp] and off == lays[q].b_recv_index(m, f, rp, hemi)


def test_bench_plane_schedule_does_exactly_n_planes_of_work():
    """bench.py pairs consecutive planes for the two-shell passes; a timed run of n planes must contain n planes of SHT work:
    floor(n/2) pair solves + (n mod 2) single solves, never a pair that reaches outside the run"""
    import bench
    for first in (0, 3, 8):
        for n in (1, 2, 5, 6, 20):
            sch = bench.plane_schedule(first, n, 2, prefetch=True)
            assert [s for s, _, _ in sch] == list(range(first, first + n))
            solved = []
            for s, partner, ahead in sch:
                if partner is not None:
                    assert partner == s + 1 and partner < first + n
                    solved += [s, partner]
                elif s not in solved:
                    solved.append(s)
                assert all(q > s for q in ahead) and len(ahead) <= 2
            assert solved == list(range(first, first + n))
            assert sum(1 for _, p, _ in sch if p is not None) == n // 2
            one = bench.plane_schedule(first, n, 1, prefetch=True)
            assert all(p is None for _, p, _ in one) and all(a == [s + 1] for s, _, a in one)


def test_b_layout_is_a_contiguous_run_in_m_for_every_ring_and_field():
    """the property the ring FFT relies on: inside a peer block the b_m of one (ring pair, field) are
    consecutive in m with the two hemispheres interleaved, so a ring's FFT streams its input instead of gathering it"""
    order, lmax = 3, 12
    nside = 1 << order
    for nranks in (1, 3):
        rp_owner, m_owner = clb.default_owners(order, lmax, nranks)
        for r in range(nranks):
            L = layout.ExchangeLayout(nside, lmax, nranks, r, rp_owner, m_owner)
            for rp in L.my_rp[:3]:
                for f in (0, 5):
                    for q in range(nranks):
                        ms = [m for m in range(lmax + 1) if m_owner[m] == q]
                        idx = [L.b_recv_index(m, f, rp, 0) for m in ms]
                        assert all(b - a == 2 for a, b in zip(idx, idx[1:]))
                        assert all(L.b_recv_index(m, f, rp, 1) == i + 1 for m, i in zip(ms, idx))
            # and on the sending side the (north, south) pair of a field is 2 consecutive elements, fields nm_mine * 2 apart
            for m in L.my_m[:2]:
                for rp in (0, 2 * nside - 1):
                    a = L.b_send_index(m, 0, rp, 0)
                    assert L.b_send_index(m, 0, rp, 1) == a + 1
                    assert L.b_send_index(m, 1, rp, 0) == a + 2 * L.my_m.size
