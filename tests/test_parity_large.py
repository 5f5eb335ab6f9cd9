"""Oracle parity at the sizes the benchmark is quoted on (-m gpu; minutes of host CPU for the reference side).

 * BASELINE configs[1] (Nside 1024, rays Nside 2048): the reference's own plane loop, run on the host cores through the
   shared-memory MPI stub (tests/test_oracle_mpi.py holds the multi-rank reference bit-identical to one rank), against
   the CUDA solver -- full comparison of every ray; plus the two transforms at lmax = 2048.
 * Nside 4096 (configs[2]), lmax = 8192 and the reference's own lmax = 3 Nside - 1 = 12287: the SAMPLED oracle
   (oracle/ref_harness.c ref_sample_*: the reference's ring FFT / plmgen / ring_synthesis, glue pinned bit-identical to
   the full functions) on ~32 m values spread over [0, lmax] and on polar / cap-boundary / equatorial ring pairs.
Tolerances as everywhere: alm <= 1e-10 relative L2, rays <= 1e-8, float maps bit-compared (<= 1e-5 of the pixels may sit
one ulp off)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref, mpirun                      # noqa: E402
from tests import mpi_workers                       # noqa: E402
from tests.test_gpu_parity import alm_err, assert_maps_match, assert_rays_match, ALM_TOL   # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


def _cores():
    n = os.cpu_count() or 1
    p = 1
    while 2 * p <= min(n, 32):
        p *= 2
    return p


def _lognormal_counts(order, seed):
    rng = np.random.default_rng(seed)
    return (8.0 * rng.lognormal(sigma=0.5, size=12 << (2 * order))).astype(np.float32)


def test_config1_transforms_lmax2048_vs_reference():
    """map2alm_mpi + filter + alm2allmaps_mpi at Nside 1024 / lmax 2048 (BASELINE configs[1]): the reference on all host
    cores (its own MPI transposes) against the CUDA library, every alm and every pixel."""
    import calclens_b200 as clb
    order, lmax = 10, 2048
    m = ((_lognormal_counts(order, 1) * np.float32(3e-4)) - np.float32(8 * np.exp(0.125) * 3e-4)).astype(np.float32)
    res = mpirun.run(_cores(), mpi_workers.sht_roundtrip, order, lmax, m, timeout=1500)
    fre = np.concatenate([r[1] for r in res]); fim = np.concatenate([r[2] for r in res])
    maps = np.zeros((6, m.size), dtype=np.float32)
    for r in res:
        maps += r[3]
    plan = clb.HEALPixSHTPlan(order, lmax)
    gre, gim = clb.map2alm_mpi(m, plan, poisson_filter=True)
    assert alm_err(gre, gim, fre, fim) <= ALM_TOL
    mg = clb.alm2allmaps_mpi(fre, fim, plan)
    assert_maps_match(mg, maps, "config[1] lmax 2048")
    plan.destroy()


def test_config1_plane_loop_vs_reference(tmp_path):
    """Two lens planes of BASELINE configs[1] (Nside 1024, full-sky rays at Nside 2048 = 5.0e7 rays, the reference's own
    lmax = 3071): the reference's plane loop (ref_driver_*: raw-map input, its map shuffles, shearinterp_comp,
    rayprop_sphere) on the host cores vs clb_solver_step on the GPU.  Every ray compared."""
    from calclens_b200 import poisson
    order, ray_order, bundle_order, nplanes = 10, 11, 5, 2
    for p in range(nplanes):
        _lognormal_counts(order, 50 + p).tofile(str(tmp_path / ("lensmap.%d" % p)))
    cosmo = poisson.Cosmology(0.27)
    max_dist = 30.0 * 20
    planes = []
    for p in range(nplanes):
        pp = poisson.plane_params(p, 20, max_dist, 0.27, cosmo)
        planes.append(dict(plane=p, wpm1=pp["wpm1"], wp=pp["wp"], wpp1=pp["wpp1"], densfact=pp["densfact"], backdens=pp["backdens"]))
    binL = max_dist / 20
    vshell = 4.0 * np.pi / 3.0 * ((planes[0]["wp"] + binL / 2) ** 3 - (planes[0]["wp"] - binL / 2) ** 3)
    part_mass = 0.27 * poisson.RHO_CRIT * vshell / (8.0 * (12 << (2 * order)))
    cfg = dict(bundle_order=bundle_order, ray_order=ray_order, map_order=order, map_path=str(tmp_path), map_name="lensmap",
               part_mass=part_mass, max_comv_distance=max_dist, num_planes=20, omega_m=0.27)
    res = mpirun.run(_cores(), mpi_workers.driver_planes_to_file, cfg, planes, str(tmp_path), timeout=3000)
    assert sum(n for _, n in res) == 12 << (2 * ray_order)
    import torch
    s = poisson.LensPlaneSolver(order, 3 * (1 << order) - 1, ray_order)
    s.init_rays(binL / 2.0)
    for p in planes:
        counts = torch.from_numpy(np.fromfile(str(tmp_path / ("lensmap.%d" % p["plane"])), dtype=np.float32)).pin_memory()
        sc = poisson.density_scalings(order, part_mass, p["densfact"], p["backdens"])
        s.step(counts, *sc, p["wpp1"], p["wp"], p["wpm1"])
    got = s.rays_host()           # NEST order from pixel 0: ray of pixel p is got[p]
    s.close()
    assert got.size == 12 << (2 * ray_order)
    # Every ray, every field.  Relative L2 <= 1e-8 (the north_star tolerance; measured ~1e-11).  Pointwise: at 7.5e7 map
    # pixels a few of the float32 map values differ from the reference's by one float ulp (FP64 round-off falling on a float
    # rounding boundary: <= 1e-5 of the pixels, see assert_maps_match), and a ray whose stencil touches such a pixel inherits
    # up to one float ulp (6e-8) of that quantity -- so pointwise the bound is one float ulp, and the rays above 1e-8 must be
    # that rare.
    fields = ("n", "beta", "A", "Aprev", "alpha", "U", "phi")
    num = {f: 0.0 for f in fields}; den = {f: 0.0 for f in fields}; mx = {f: 0.0 for f in fields}; sc = {f: 0.0 for f in fields}
    chunks = []
    for fn, n in res:             # rank by rank (the reference's ranks own Peano ranges of bundle cells)
        want = np.load(fn, mmap_mode="r")
        for lo in range(0, n, 1 << 22):
            w = np.array(want[lo:lo + (1 << 22)])
            g = got[w["nest"]]
            assert np.array_equal(g["nest"], w["nest"])
            for f in fields:
                d = np.abs(g[f] - w[f])
                num[f] += float((d ** 2).sum()); den[f] += float((w[f] ** 2).sum())
                sc[f] = max(sc[f], float(np.abs(w[f]).max()))
                chunks.append((f, d.reshape(d.shape[0], -1).max(axis=1)))
    nrays = 12 << (2 * ray_order)
    for f in fields:
        rel_l2 = np.sqrt(num[f] / max(den[f], 1e-300))
        dmax = max(float(c.max()) for ff, c in chunks if ff == f)
        above = sum(int((c > 1e-8 * sc[f]).sum()) for ff, c in chunks if ff == f)
        print("config[1] plane loop, ray field %-5s: rel L2 %.2e, max |diff| / max |ref| %.2e, rays above 1e-8: %d of %d" % (
            f, rel_l2, dmax / max(sc[f], 1e-300), above, nrays))
        assert rel_l2 <= 1e-8, (f, rel_l2)
        assert dmax <= 2e-7 * sc[f], (f, dmax, sc[f])
        assert above <= 1e-4 * nrays, (f, above)


def _sample_m(lmax):
    base = [0, 1, 2, 3, 39, 40, 41, 50, 59, 60, 61, 127, 128, 1000, 1023, 1024, 2047, 2048, 4095, 4096, 4097,
            lmax // 2, 6000, 8000, 8191, lmax - 2048, lmax - 100, lmax - 2, lmax - 1, lmax]
    extra = np.random.default_rng(lmax).integers(0, lmax + 1, 6).tolist()
    return sorted({int(x) for x in base + extra if 0 <= x <= lmax})


@pytest.mark.parametrize("lmax", [8192, 12287])
def test_nside4096_sampled_oracle(lmax):
    """Nside 4096 at the quoted band limit (8192) and at the one the reference actually runs (3 Nside - 1 = 12287):
    analysis of a lognormal shell checked on ~35 m values (all l), synthesis checked pixel by pixel on 14 ring pairs
    (polar cap, cap boundary, equatorial belt, equator; both hemispheres; all six maps)."""
    import torch
    import calclens_b200 as clb
    order = 12
    nside = 1 << order
    m = ((_lognormal_counts(order, 7) * np.float32(3e-4)) - np.float32(8 * np.exp(0.125) * 3e-4)).astype(np.float32)
    ms = _sample_m(lmax)
    rows = ref.sample_map2alm(order, lmax, m, ms)
    plan = clb.HEALPixSHTPlan(order, lmax)
    dm = torch.from_numpy(m).cuda()
    are, aim = plan.legendre_analysis(plan.ring_analysis(dm), poisson_filter=False)
    are_h = are.cpu().numpy(); aim_h = aim.cpu().numpy()
    worst = 0.0
    for mm in ms:
        off = clb.lm2index(mm, mm, lmax)
        k = lmax - mm + 1
        e = alm_err(are_h[off:off + k], aim_h[off:off + k], rows[mm][0], rows[mm][1])
        worst = max(worst, e)
        assert e <= ALM_TOL, "m = %d: alm relative L2 error %g" % (mm, e)
    # synthesis from a red spectrum (the filtered shell), six maps on sampled ring pairs
    fre, fim = plan.legendre_analysis(plan.ring_analysis(dm), poisson_filter=True)
    fre_h = fre.cpu().numpy(); fim_h = fim.cpu().numpy()
    maps = plan.ring_synthesis(plan.legendre_synthesis(fre, fim))
    rings = [1, 2, 3, 7, 100, 1365, 2731, nside - 1, nside, nside + 1, nside + 2, 6001, 2 * nside - 1, 2 * nside]
    samp = ref.sample_alm2allmaps_rings(order, lmax, fre_h, fim_h, rings)
    npix = 12 * nside * nside
    for ring in rings:
        n = 4 * ring if ring < nside else 4 * nside
        s0 = 2 * ring * (ring - 1) if ring < nside else 2 * nside * (nside - 1) + (ring - nside) * 4 * nside
        north, south = samp[ring]
        assert_maps_match(maps[:, s0:s0 + n].cpu().numpy(), north, "lmax %d ring %d north" % (lmax, ring))
        if south is not None:
            s1 = npix - s0 - n
            assert_maps_match(maps[:, s1:s1 + n].cpu().numpy(), south, "lmax %d ring %d south" % (lmax, ring))
    plan.destroy()
    print("Nside 4096 lmax %d: worst sampled-m alm error %.2e over %d m values; %d ring pairs bit-compared" % (lmax, worst, len(ms), len(rings)))
