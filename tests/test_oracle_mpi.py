"""CPU-side checks of the oracle infrastructure added in round 2 (no GPU):
 * the sampled oracle (reference primitives + restated glue, used at Nside 4096) equals the full reference functions
   bit for bit where both can run;
 * the reference's MPI code run on several ranks of the shared-memory MPI stub (real hypercube transposes,
   map2alm_transpose_mpi.c:356-381, alm2allmaps_transpose_mpi.c:699-724, and the ring<->domain shuffles of map_shuffle.c)
   reproduces its single-rank results;
 * the reference's own plane loop (ref_driver_*) agrees with the function-by-function oracle chain."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref, mpirun   # noqa: E402
from tests import mpi_workers    # noqa: E402

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def _counts(order, seed):
    rng = np.random.default_rng(seed)
    return (8.0 * rng.lognormal(sigma=0.5, size=12 << (2 * order))).astype(np.float32)


@pytest.mark.parametrize("order,lmax", [(4, 47), (4, 32), (5, 64)])
def test_sampled_oracle_equals_full_reference(order, lmax):
    m = (_counts(order, 3) - np.float32(8.0)).astype(np.float32)
    w = np.random.default_rng(1).normal(scale=1e-3, size=2 << order)
    are, aim = ref.map2alm(order, lmax, m, ring_weights=w)
    rows = ref.sample_map2alm(order, lmax, m, range(lmax + 1), ring_weights=w)
    off = 0
    for mm in range(lmax + 1):
        k = lmax - mm + 1
        assert np.array_equal(rows[mm][0], are[off:off + k]) and np.array_equal(rows[mm][1], aim[off:off + k]), mm
        off += k
    fre, fim = ref.poisson_filter(lmax, are, aim)
    maps = ref.alm2allmaps(order, lmax, fre, fim)
    nside = 1 << order
    npix = 12 * nside * nside
    samp = ref.sample_alm2allmaps_rings(order, lmax, fre, fim, range(1, 2 * nside + 1))
    L = ref.lib()
    for ring in range(1, 2 * nside + 1):
        import ctypes as C
        sp, rp, sh = C.c_long(), C.c_long(), C.c_long(); ct, st = C.c_double(), C.c_double()
        L.get_ring_info2(ring, C.byref(sp), C.byref(rp), C.byref(ct), C.byref(st), C.byref(sh), order)
        n, s0 = rp.value, sp.value
        north, south = samp[ring]
        assert np.array_equal(north.view(np.uint32), maps[:, s0:s0 + n].view(np.uint32)), ring
        if south is not None:
            s1 = npix - s0 - n
            assert np.array_equal(south.view(np.uint32), maps[:, s1:s1 + n].view(np.uint32)), ring


@pytest.mark.parametrize("ntasks", [2, 3])
def test_shm_mpi_reference_matches_single_rank(ntasks):
    order, lmax = 4, 47     # the reference's own lmax = 3 Nside - 1, its own ring / m split per rank
    m = (_counts(order, 5) - np.float32(8.0)).astype(np.float32)
    are, aim = ref.map2alm(order, lmax, m)
    fre, fim = ref.poisson_filter(lmax, are, aim)
    maps = ref.alm2allmaps(order, lmax, fre, fim)
    res = mpirun.run(ntasks, mpi_workers.sht_roundtrip, order, lmax, m, timeout=300)
    got_re = np.concatenate([r[1] for r in res]); got_im = np.concatenate([r[2] for r in res])
    assert [r[0]["first_m"] for r in res] == sorted(r[0]["first_m"] for r in res)
    assert np.array_equal(got_re, fre) and np.array_equal(got_im, fim)
    tot = np.zeros_like(maps)
    for r in res:
        tot += r[3]            # disjoint ring sets
    assert np.array_equal(tot.view(np.uint32), maps.view(np.uint32))


def _driver_cfg(tmp_path, order, ray_order, bundle_order, nplanes):
    from calclens_b200 import poisson
    counts = _counts(order, 11)
    counts.tofile(str(tmp_path / "lensmap.0"))
    _counts(order, 12).tofile(str(tmp_path / "lensmap.1"))
    cosmo = poisson.Cosmology(0.27)
    max_dist = 30.0 * nplanes
    planes = []
    for p in range(nplanes):
        pp = poisson.plane_params(p, nplanes, max_dist, 0.27, cosmo)
        planes.append(dict(plane=p, wpm1=pp["wpm1"], wp=pp["wp"], wpp1=pp["wpp1"], densfact=pp["densfact"] * 1e3, backdens=pp["backdens"] * 1e3))
    cfg = dict(bundle_order=bundle_order, ray_order=ray_order, map_order=order, map_path=str(tmp_path), map_name="lensmap",
               part_mass=3.0e10, max_comv_distance=max_dist, num_planes=nplanes, omega_m=0.27)
    return cfg, planes


def test_reference_plane_loop_matches_function_chain(tmp_path):
    """ref_driver_* (the reference's do_healpix_sht_poisson_solve with its map shuffles + rayprop_sphere loop) against the
    function-by-function chain the other tests use (map2alm_mpi -> filter -> alm2allmaps_mpi -> shearinterp_comp ->
    rayprop_sphere on full-sky maps).  lmax is the reference's own 3 Nside - 1 here."""
    from calclens_b200 import poisson
    order, ray_order, bundle_order, nplanes = 4, 4, 1, 2
    cfg, planes = _driver_cfg(tmp_path, order, ray_order, bundle_order, nplanes)
    rays_d = mpi_workers.driver_planes(0, 1, cfg, planes)
    rays_d = rays_d[np.argsort(rays_d["nest"])]
    lmax = 3 * (1 << order) - 1
    rays = ref.init_rays(ray_order, cfg["max_comv_distance"] / nplanes / 2.0)
    for p in planes:
        counts = np.fromfile(os.path.join(cfg["map_path"], "lensmap.%d" % p["plane"]), dtype=np.float32)
        pm, dm, bd = poisson.density_scalings(order, cfg["part_mass"], p["densfact"], p["backdens"])
        dens = ((counts * pm) * dm - bd).astype(np.float32)
        are, aim = ref.map2alm(order, lmax, dens)
        are, aim = ref.poisson_filter(lmax, are, aim)
        maps = ref.alm2allmaps(order, lmax, are, aim)
        for f in ("phi", "alpha", "U"):
            rays[f] = 0.0
        ref.shearinterp(order, bundle_order, maps, rays)
        ref.rayprop(rays, p["wpp1"], p["wp"], p["wpm1"])
    assert np.array_equal(rays_d["nest"], rays["nest"])
    for f in ("n", "beta", "A", "Aprev", "alpha", "U", "phi"):
        assert np.array_equal(rays_d[f], rays[f]), f


def test_reference_plane_loop_two_ranks(tmp_path):
    """the same plane loop on 2 ranks of the shared-memory MPI stub (domain decomposition, peano2ring / ring2peano shuffles
    with halo cells, hypercube transposes) gives the single-rank rays bit for bit"""
    order, ray_order, bundle_order, nplanes = 4, 4, 1, 2
    cfg, planes = _driver_cfg(tmp_path, order, ray_order, bundle_order, nplanes)
    one = mpi_workers.driver_planes(0, 1, cfg, planes)
    one = one[np.argsort(one["nest"])]
    res = mpirun.run(2, mpi_workers.driver_planes, cfg, planes, timeout=300)
    two = np.concatenate(res)
    two = two[np.argsort(two["nest"])]
    assert two.size == one.size and np.array_equal(two["nest"], one["nest"])
    for f in ("n", "beta", "A", "Aprev", "alpha", "U", "phi"):
        assert np.array_equal(two[f], one[f]), f
