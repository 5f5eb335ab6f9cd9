"""Link-level drop-in test (-m gpu).  oracle/Makefile builds the SAME harness twice: against the reference's own
map2alm_transpose_mpi.o / alm2allmaps_transpose_mpi.o / rayprop.o / do_healpix_sht_poisson_solve (libcalclens_ref.so),
and against shim/calclens_b200_shim.c, which defines those four symbols with the reference's prototypes and forwards
them to libcalclens_b200.so (libcalclens_ref_shim.so).  The same harness calls then run through both libraries:
everything around the replaced functions -- healpixsht_plan, the plan arrays, mapvec packing, bundle cells, ray
allocation, the plane loop -- is the reference's own host code in both cases."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref, mpirun                      # noqa: E402
from tests import mpi_workers                       # noqa: E402
from tests.test_gpu_parity import alm_err, assert_maps_match, assert_rays_match, ALM_TOL   # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (ref.available() and ref.available("shim")), reason="oracle/_ref (shim) not built")]


def _map(order, seed):
    rng = np.random.default_rng(seed)
    npix = 12 << (2 * order)
    return ((8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32) * np.float32(3e-4) - np.float32(8 * np.exp(0.125) * 3e-4)).astype(np.float32)


def test_shim_library_defines_the_reference_symbols():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", ref.path("shim")], capture_output=True, text=True).stdout
    syms = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for s in ("map2alm_mpi", "alm2allmaps_mpi", "rayprop_sphere", "do_healpix_sht_poisson_solve", "calclens_b200_rayprop_all"):
        assert s in syms, s
    assert ref.lib("shim").ref_is_shim() == 1 and ref.lib("ref").ref_is_shim() == 0


@pytest.mark.parametrize("order,lmax,use_w", [(3, 23, False), (4, 32, True), (5, 95, True), (6, 128, False), (7, 383, True)])
def test_shim_transforms_match_reference(order, lmax, use_w):
    """map2alm_mpi / alm2allmaps_mpi through the shim (GPU) vs the reference objects, same harness, lmax = 2 Nside and the
    reference's own 3 Nside - 1, with and without ring weights"""
    w = None
    if use_w:
        w = np.load(os.path.join(ROOT, "tests", "golden", "ring_weights.npz"))["n%05d" % (1 << order)]
    m = _map(order, 300 + order)
    are, aim = ref.map2alm(order, lmax, m, w, variant="ref")
    gre, gim = ref.map2alm(order, lmax, m, w, variant="shim")
    assert alm_err(gre, gim, are, aim) <= ALM_TOL
    fre, fim = ref.poisson_filter(lmax, are, aim)
    mo = ref.alm2allmaps(order, lmax, fre, fim, variant="ref")
    mg = ref.alm2allmaps(order, lmax, fre, fim, variant="shim")
    assert_maps_match(mg, mo, "shim order %d lmax %d" % (order, lmax))


def test_shim_rayprop_sphere_matches_reference():
    rng = np.random.default_rng(4)
    a = ref.init_rays(6, 15.0)[::3].copy()
    a["alpha"] = rng.normal(scale=2e-4, size=a["alpha"].shape)
    a["U"] = rng.normal(scale=3e-3, size=a["U"].shape)
    a[:50]["alpha"] = 0.0                       # the undeflected branch (rayprop.c:124-131)
    b = a.copy()
    ref.rayprop(a, 75.0, 45.0, 15.0, variant="ref")
    ref.rayprop(b, 75.0, 45.0, 15.0, variant="shim")
    assert_rays_match(b, a)


def _plane_setup(tmp_path, order, ray_order, bundle_order, nplanes):
    from tests.test_oracle_mpi import _driver_cfg
    return _driver_cfg(tmp_path, order, ray_order, bundle_order, nplanes)


@pytest.mark.parametrize("resident", [0, 1])
def test_shim_plane_loop_matches_reference(tmp_path, resident):
    """The reference's plane loop (domain decomposition, ray allocation, ray reset, do_healpix_sht_poisson_solve,
    rayprop_sphere per bundle cell) with the four symbols replaced by the shim: coarse entry on the GPU, rays either
    round-tripping every call (default) or device resident (CALCLENS_B200_RESIDENT=1)."""
    order, ray_order, bundle_order, nplanes = 5, 6, 2, 2
    cfg, planes = _plane_setup(tmp_path, order, ray_order, bundle_order, nplanes)
    want = mpi_workers.driver_planes(0, 1, cfg, planes, "ref")
    got = mpirun.run(1, mpi_workers.driver_planes, cfg, planes, "shim", timeout=600,
                     extra_env={"CALCLENS_B200_RESIDENT": str(resident)})[0]
    assert np.array_equal(got["nest"], want["nest"])
    assert_rays_match(got, want)


@pytest.mark.parametrize("ntasks", [2, 3])
def test_shim_multi_rank_transforms(ntasks):
    """NTasks > 1: every rank passes the reference's per-rank plan (firstRingTasks/lastRingTasks, firstMTasks/lastMTasks,
    its own mapvec slice) to the shim; the transposes run through the fused exchange (ranks share the one GPU of the box:
    CUDA IPC between processes, host barriers).  Against the single-rank reference."""
    order, lmax = 5, 95
    m = _map(order, 41)
    are, aim = ref.map2alm(order, lmax, m)
    fre, fim = ref.poisson_filter(lmax, are, aim)
    maps = ref.alm2allmaps(order, lmax, fre, fim)
    res = mpirun.run(ntasks, mpi_workers.sht_roundtrip, order, lmax, m, "shim", timeout=600)
    got_re = np.concatenate([r[1] for r in res]); got_im = np.concatenate([r[2] for r in res])
    assert alm_err(got_re, got_im, fre, fim) <= ALM_TOL
    tot = np.zeros_like(maps)
    for r in res:
        tot += r[3]
    # (each rank synthesised from ITS alm slice of the GPU analysis, not from the reference's alm: FP64-level input noise)
    assert_maps_match(tot, maps, "shim %d ranks" % ntasks)


def test_shim_multi_rank_plane_loop(tmp_path):
    """the plane loop on 2 MPI ranks through the shim (coarse entry, full-map broadcast to both ranks) vs the reference on
    one rank"""
    order, ray_order, bundle_order, nplanes = 5, 5, 2, 2
    cfg, planes = _plane_setup(tmp_path, order, ray_order, bundle_order, nplanes)
    want = mpi_workers.driver_planes(0, 1, cfg, planes, "ref")
    want = want[np.argsort(want["nest"])]
    res = mpirun.run(2, mpi_workers.driver_planes, cfg, planes, "shim", timeout=600)
    got = np.concatenate(res)
    got = got[np.argsort(got["nest"])]
    assert np.array_equal(got["nest"], want["nest"])
    assert_rays_match(got, want)
