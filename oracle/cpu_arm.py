"""oracle/cpu_arm.py -- TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's reference arm and cpu_baseline leg).

Times the reference's own CPU implementation of the per-plane hot path on the host cores: N ranks of the shared-memory
MPI stub (oracle/stubs/mpi_stub.c) run the UNMODIFIED map2alm_mpi / alm2allmaps_mpi (real hypercube transposes) on the
reference's own ring / m decomposition, then every rank interpolates and propagates its share of the rays
(shearinterp_comp + rayprop_sphere).  Library: oracle/_ref/libcalclens_ref_fast.so = the reference built with its own
production flags (-O3 -ffast-math -funroll-loops, /root/reference/Makefile:74); the parity build is used if it is absent.
"""
import math
import os
import time

import numpy as np


def variant():
    from . import ref
    return "fast" if ref.available("fast") else "ref"


def _shared_maps(path, npix, create):
    return np.memmap(path, dtype=np.float32, mode="w+" if create else "r+", shape=(6, npix))


def plane_worker(rank, ntasks, order, lmax, ray_order, seed, shared_path, var):
    """one lens plane at (order, lmax), rays at ray_order split evenly over the ranks; returns the stage seconds"""
    from . import ref
    L = ref.lib(var)
    nside = 1 << order
    npix = 12 * nside * nside
    rng = np.random.default_rng(seed)
    m = rng.lognormal(sigma=0.5, size=npix).astype(np.float32)
    m = (m * np.float32(8.0) - np.float32(8.0 * math.exp(0.125))).astype(np.float32)
    L.MPI_Barrier(0)
    t0 = time.time()
    info, are, aim = ref.mpi_map2alm(order, lmax, m, variant=var)
    L.MPI_Barrier(0)
    t1 = time.time()
    k = 0
    for mm in range(info["first_m"], info["last_m"] + 1):       # shtpoissonsolve.c:526-550 on the local slice
        n = lmax - mm + 1
        l = np.arange(mm, lmax + 1, dtype=np.float64)
        f = np.where(l > 0, -1.0 / np.maximum(l, 1.0) / (l + 1.0), 0.0)
        are[k:k + n] *= f; aim[k:k + n] *= f
        k += n
    maps = ref.mpi_alm2allmaps(order, lmax, are, aim, variant=var)
    L.MPI_Barrier(0)
    t2 = time.time()
    # every rank needs the maps under its rays: own rings go to a shared file (stands in for the ring -> domain shuffle,
    # map_shuffle.c, which the plane-loop variant of this arm runs for real; its time is reported separately there)
    sh = _shared_maps(shared_path, npix, False)
    first, last = info["first_ring"], info["last_ring"]
    lo = 2 * first * (first - 1) if first <= nside else 2 * nside * (nside - 1) + (first - nside) * 4 * nside
    r1 = last + 1
    hi = 2 * r1 * (r1 - 1) if r1 <= nside else 2 * nside * (nside - 1) + (r1 - nside) * 4 * nside
    hi = min(hi, npix - lo) if last == 2 * nside else hi
    sh[:, lo:hi] = maps[:, lo:hi]
    sh[:, npix - hi:npix - lo] = maps[:, npix - hi:npix - lo]
    sh.flush()
    L.MPI_Barrier(0)
    full = np.array(sh)
    full *= np.float32(1e-3 / max(float(np.abs(full[3]).max()), 1e-30))
    nrays = 12 << (2 * ray_order)
    a, b = (nrays * rank) // ntasks, (nrays * (rank + 1)) // ntasks
    rays = ref.init_rays(ray_order, 15.0, first=a, n=b - a)
    L.MPI_Barrier(0)
    t3 = time.time()
    ref.shearinterp(order, min(order, 3), full, rays, variant=var)
    ref.rayprop(rays, 45.0, 15.0, 0.0, variant=var)
    L.MPI_Barrier(0)
    t4 = time.time()
    return dict(analysis=t1 - t0, synthesis=t2 - t1, rays=t4 - t3)


def fft_only_seconds(order, var):
    """single-core time of the ring FFTs of one map (analysis direction): ref_sample_map2alm with no m selected"""
    from . import ref
    rng = np.random.default_rng(0)
    m = rng.normal(size=12 << (2 * order)).astype(np.float32)
    t = time.time()
    ref.sample_map2alm(order, 2 << order, m, [], variant=var)
    return time.time() - t


def npix_log(nside):
    """sum over rings of n log2 n (FFT work model)"""
    r = np.arange(1, 2 * nside + 1, dtype=np.float64)
    n = np.where(r < nside, 4 * r, 4.0 * nside)
    w = n * np.log2(n)
    return float(2 * w[:-1].sum() + w[-1])


def sample(order, lmax, ray_order, cores, seed, triple_count, target):
    """Run one plane at the sample size on `cores` ranks and extrapolate to target = (nside, lmax, nrays).
    Returns (planes_per_s, detail dict, wall seconds)."""
    import tempfile
    from . import mpirun
    var = variant()
    nside = 1 << order
    fd, shared = tempfile.mkstemp(prefix="clb_maps_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    os.close(fd)
    try:
        _shared_maps(shared, 12 * nside * nside, True).flush()
        t0 = time.time()
        if cores > 1:
            res = mpirun.run(cores, plane_worker, order, lmax, ray_order, seed, shared, var, timeout=3000)
        else:
            res = [plane_worker(0, 1, order, lmax, ray_order, seed, shared, var)]
        wall = time.time() - t0
    finally:
        os.unlink(shared)
    t_ana = max(r["analysis"] for r in res); t_syn = max(r["synthesis"] for r in res); t_ray = max(r["rays"] for r in res)
    t_fft1 = fft_only_seconds(order, var)                 # one map, one core
    fft_share = min(7.0 * t_fft1 / cores, 0.9 * (t_ana + t_syn))   # 1 analysis + 6 synthesis maps, assumed to parallelise perfectly
    leg = t_ana + t_syn - fft_share
    t_nside, t_lmax, t_nrays = target
    f_tri = triple_count(t_nside, t_lmax) / triple_count(nside, lmax)
    f_fft = npix_log(t_nside) / npix_log(nside)
    f_ray = t_nrays / float(12 << (2 * ray_order))
    per_plane = leg * f_tri + fft_share * f_fft + t_ray * f_ray
    detail = {"library": "oracle/_ref/libcalclens_ref_%s.so" % var,
              "flags": "-O3 -ffast-math -funroll-loops (reference Makefile:74)" if var == "fast" else "-O2 -ffp-contract=off (parity build)",
              "ranks": cores, "sample": {"nside": nside, "lmax": lmax, "ray_nside": 1 << ray_order},
              "measured_seconds": {"map2alm_mpi": t_ana, "filter+alm2allmaps_mpi": t_syn, "shearinterp_comp+rayprop_sphere": t_ray,
                                   "ring_fft_one_map_one_core": t_fft1, "wall_incl_setup": wall},
              "extrapolation": {"legendre_by_triples": f_tri, "fft_by_npix_log_n": f_fft, "rays_by_count": f_ray,
                                "legendre_seconds_at_sample": leg, "fft_seconds_at_sample": fft_share},
              "extrapolated_seconds_per_plane": per_plane}
    return 1.0 / per_plane, detail, wall
