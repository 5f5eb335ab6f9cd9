"""CPU-side: pin the oracle.  (1) the compiled reference (oracle/_ref, when present) and the restatement
(oracle/port) against the committed golden vectors, which were produced by the unmodified reference functions
(tools/make_golden.py); (2) analytic identities the reference satisfies."""
import os

import numpy as np
import pytest

from tests import oracle_select

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ORACLES = oracle_select.both()
IDS = [o.NAME.split()[0] for o in ORACLES]


@pytest.fixture(scope="module")
def g_sht():
    return np.load(os.path.join(GOLD, "sht.npz"))


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_indexing_golden(orc):
    g = np.load(os.path.join(GOLD, "healpix_index.npz"))
    for order in (0, 1, 2, 3, 4, 5):
        npix = 12 << (2 * order)
        assert [orc.ring2nest(i, order) for i in range(npix)] == list(g["ring2nest_o%d" % order])
        assert [orc.nest2ring(i, order) for i in range(npix)] == list(g["nest2ring_o%d" % order])
        assert [orc.nest2peano(i, order) for i in range(npix)] == list(g["nest2peano_o%d" % order])
        assert sorted(g["ring2nest_o%d" % order]) == list(range(npix))      # a permutation
    th, ph = g["theta"], g["phi"]
    for order in (0, 3, 8, 13):
        got = [orc.ang2nest(t, p, order) for t, p in zip(th, ph)]
        assert got == list(g["ang2nest_o%d" % order])
    for order in (1, 4, 10):
        for i in range(0, th.size, 7):
            t = th[i]
            if t == 0.0 or t == np.pi:
                t = 1e-7 if t == 0.0 else np.pi - 1e-7
            pix, wgt = orc.get_interpol(t, ph[i], order)
            assert pix == list(g["interpol_pix_o%d" % order][i])
            assert np.array_equal(np.array(wgt), g["interpol_wgt_o%d" % order][i])
            assert abs(sum(wgt) - 1.0) < 1e-12


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_sht_golden_bit_exact(orc, g_sht):
    wts = np.load(os.path.join(GOLD, "ring_weights.npz"))
    for tag in "abcd":
        order, lmax = int(g_sht[tag + "_order"]), int(g_sht[tag + "_lmax"])
        w = wts["n%05d" % (1 << order)] if int(g_sht[tag + "_weights"]) else None
        are, aim = orc.map2alm(order, lmax, g_sht[tag + "_map"], w)
        assert np.array_equal(are, g_sht[tag + "_alm_re"]) and np.array_equal(aim, g_sht[tag + "_alm_im"])
        fre, fim = orc.poisson_filter(lmax, are, aim)
        assert np.array_equal(fre, g_sht[tag + "_falm_re"]) and np.array_equal(fim, g_sht[tag + "_falm_im"])
        maps = orc.alm2allmaps(order, lmax, fre, fim)
        assert np.array_equal(maps, g_sht[tag + "_maps"])


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_plm_golden(orc, g_sht):
    for (lmax, cth, m, firstl), vals in zip(g_sht["plm_args"], g_sht["plm_vals"]):
        f, vec = orc.plmgen(int(lmax), float(cth), float(np.sqrt((1 - cth) * (1 + cth))), int(m))
        assert f == int(firstl)
        assert np.array_equal(vec[f:f + 8], vals)


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_rays_golden_bit_exact(orc):
    g = np.load(os.path.join(GOLD, "rays.npz"))
    order = int(g["order"])
    rays = g["rays0"].view(orc.RAY_DTYPE).copy()
    orc.shearinterp(order, 2, g["maps"], rays)
    assert rays.tobytes() == g["rays_interp"].tobytes()
    orc.rayprop(rays, 45.0, 15.0, 0.0)
    assert rays.tobytes() == g["rays_prop1"].tobytes()
    rays["alpha"] = 0; rays["U"] = 0; rays["phi"] = 0
    orc.shearinterp(order, 2, g["maps"], rays)
    orc.rayprop(rays, 75.0, 45.0, 15.0)
    assert rays.tobytes() == g["rays_prop2"].tobytes()
    rz = g["rz0"].view(orc.RAY_DTYPE).copy()
    orc.rayprop(rz, 45.0, 15.0, 0.0)
    assert rz.tobytes() == g["rz1"].tobytes()


def test_port_equals_reference_on_fresh_inputs():
    """beyond the committed vectors: restatement == compiled reference, bit for bit, on new seeded inputs"""
    if len(ORACLES) < 2:
        pytest.skip("compiled reference (oracle/_ref) not present")
    port, ref = ORACLES
    rng = np.random.default_rng(11)
    for order, lmax in ((2, 11), (3, 16), (5, 95)):
        npix = 12 << (2 * order)
        m = (rng.lognormal(size=npix) - 1.6).astype(np.float32)
        a, b = ref.map2alm(order, lmax, m), port.map2alm(order, lmax, m)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        f = ref.poisson_filter(lmax, *a)
        assert np.array_equal(ref.alm2allmaps(order, lmax, *f), port.alm2allmaps(order, lmax, *f))


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_identities(orc):
    """single (l,m): grad_tt + grad_pp = -l(l+1) phi to float precision, and analysis(synthesis) returns the alm
    to HEALPix quadrature accuracy."""
    order, lmax = 4, 47
    l0, m0 = 5, 2
    idx = m0 * (lmax + 1) - m0 * (m0 - 1) // 2 + (l0 - m0)
    n = (lmax + 1) * (lmax + 2) // 2
    are = np.zeros(n); aim = np.zeros(n)
    are[idx], aim[idx] = 1.0, 0.5
    maps = orc.alm2allmaps(order, lmax, are, aim)
    scale = np.abs(maps[0]).max() * l0 * (l0 + 1)
    assert np.abs(maps[3] + maps[5] + l0 * (l0 + 1) * maps[0]).max() < 1e-5 * scale
    wts = np.load(os.path.join(GOLD, "ring_weights.npz"))["n%05d" % (1 << order)]
    bre, bim = orc.map2alm(order, lmax, maps[0], wts)
    assert abs(bre[idx] - 1.0) < 1e-4 and abs(bim[idx] - 0.5) < 1e-4
    assert np.sqrt(((bre - are) ** 2 + (bim - aim) ** 2).sum()) < 1e-2


def test_point_mass_sign_and_scale(g_sht):
    """|grad phi|(d) ~ (S/4pi) cot(d/2) around a unit point mass: sign, units, normalisation
    of the whole reference chain (unsmoothed single pixel, so only to ~15%)."""
    from oracle import port
    order = 4
    npix = 12 << (2 * order)
    maps = g_sht["pm_maps"]
    pix = int(g_sht["pm_pixel"])
    vc = np.zeros(3); port.lib().port_nest2vec(port.ring2nest(pix, order), vc.ctypes.data, order)
    S = 1.0   # the map integrates to S * (pixel area)^-1 * area = 1 per unit pixel area -> S = pixel area
    S = 4 * np.pi / npix
    ratios = []
    for p in range(0, npix, 37):
        v = np.zeros(3); port.lib().port_nest2vec(port.ring2nest(p, order), v.ctypes.data, order)
        d = np.arccos(np.clip(v @ vc, -1, 1))
        if 0.3 < d < 2.0:
            g = np.hypot(maps[1][p], maps[2][p])
            ratios.append(g / (S / (4 * np.pi) / np.tan(d / 2)))
    assert 0.85 < np.mean(ratios) < 1.15


def test_port_matches_reference_on_next_rows():
    """The restatement of the two 'next' rows (ray output transform, NGP deposit) is bit-identical to the reference
    functions compiled in oracle/_ref (paratrans_ray_curr2obs, rot_ray_ang2radec, vec2ang + ang2nest)."""
    from oracle import port, ref
    if not ref.available():
        import pytest
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1)
    order = 5
    pos = rng.normal(size=(20000, 3)).astype(np.float32)
    pos[:3] = [[0, 0, 1], [0, 0, -1], [1, 0, 0]]
    mass = (3.7e9 * (0.5 + rng.random(20000))).astype(np.float32)
    assert np.array_equal(ref.deposit_ngp(pos, mass, order), port.deposit_ngp(pos, mass, order))
    rays = ref.init_rays(4, 15.0)
    rays["n"] += rng.normal(size=rays["n"].shape) * 0.01
    for f, n in (("A", 4), ("Aprev", 4), ("U", 4), ("alpha", 2)):
        rays[f] += rng.normal(size=(rays.size, n)) * 0.05
    r1 = rays.copy(); r2 = rays.copy()
    ref.ray_output(r1, 4); port.ray_output(r2, 4)
    assert r1.tobytes() == r2.tobytes()
    r1 = rays.copy(); r2 = rays.copy()
    ref.rayprop_born(r1, 45.0, 15.0, 0.0); port.rayprop_born(r2, 45.0, 15.0, 0.0)
    ref.rayprop_born(r1, 75.0, 45.0, 15.0); port.rayprop_born(r2, 75.0, 45.0, 15.0)
    assert r1.tobytes() == r2.tobytes()


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_next_rows_golden(orc):
    """NGP deposit, write_rays' output transform and the Born step against vectors produced by the unmodified reference
    functions (tools/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "next_rows.npz"))
    assert np.array_equal(orc.deposit_ngp(g["dep_pos"], g["dep_mass"], int(g["dep_order"])), g["dep_map"])
    rays = g["rays_in"].copy().view(orc.RAY_DTYPE)
    r = rays.copy(); orc.ray_output(r, 3)
    assert r.tobytes() == g["rays_output"].tobytes()
    r = rays.copy(); orc.rayprop_born(r, 45.0, 15.0, 0.0); orc.rayprop_born(r, 75.0, 45.0, 15.0)
    assert r.tobytes() == g["rays_born"].tobytes()
