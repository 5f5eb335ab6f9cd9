"""Module-level worker functions for oracle/mpirun.py (they must be importable in spawned processes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def sht_roundtrip(rank, ntasks, order, lmax, ringmap, variant="ref"):
    """every rank: map2alm_mpi -> Poisson filter on its slice -> alm2allmaps_mpi; returns (plan info, alm, own rings of the maps)"""
    from oracle import ref
    info, are, aim = ref.mpi_map2alm(order, lmax, ringmap, variant=variant)
    # filter on the local slice (shtpoissonsolve.c:526-550)
    k = 0
    are = are.copy(); aim = aim.copy()
    for m in range(info["first_m"], info["last_m"] + 1):
        for l in range(m, lmax + 1):
            if l == 0:
                are[k] = 0.0; aim[k] = 0.0
            else:
                f = -1.0 / float(l) / (float(l) + 1.0)
                are[k] *= f; aim[k] *= f
            k += 1
    maps = ref.mpi_alm2allmaps(order, lmax, are, aim, variant=variant)
    return info, are, aim, maps


def driver_planes(rank, ntasks, cfg, planes, variant="ref"):
    """the reference's plane loop (ref_driver_*) on this rank; returns this rank's rays after the last plane"""
    from oracle import ref
    ref.driver_init(cfg["bundle_order"], cfg["ray_order"], cfg["map_order"], cfg["map_path"], cfg["map_name"], cfg["part_mass"],
                    cfg["max_comv_distance"], cfg["num_planes"], cfg["omega_m"], cfg.get("ring_weight_path", ""), variant=variant)
    for p in planes:
        ref.driver_plane(p["plane"], p["wpm1"], p["wp"], p["wpp1"], p["densfact"], p["backdens"], variant=variant)
    rays = ref.driver_rays(variant=variant)
    ref.driver_finalize(variant=variant)
    return rays


def driver_planes_to_file(rank, ntasks, cfg, planes, out_dir, variant="ref"):
    """driver_planes for large runs: this rank's rays go to <out_dir>/rays.<rank>.npy instead of through the result queue"""
    rays = driver_planes(rank, ntasks, cfg, planes, variant)
    fn = os.path.join(out_dir, "rays.%d.npy" % rank)
    np.save(fn, rays)
    return fn, int(rays.size)


def timed_driver_planes(rank, ntasks, cfg, planes, variant="fast"):
    """as driver_planes, timing every plane (seconds; the caller takes the max over ranks)"""
    import time
    from oracle import ref
    ref.driver_init(cfg["bundle_order"], cfg["ray_order"], cfg["map_order"], cfg["map_path"], cfg["map_name"], cfg["part_mass"],
                    cfg["max_comv_distance"], cfg["num_planes"], cfg["omega_m"], cfg.get("ring_weight_path", ""), variant=variant)
    L = ref.lib(variant)
    times = []
    for p in planes:
        L.MPI_Barrier(0)
        t0 = time.time()
        ref.driver_plane(p["plane"], p["wpm1"], p["wp"], p["wpp1"], p["densfact"], p["backdens"], variant=variant)
        L.MPI_Barrier(0)
        times.append(time.time() - t0)
    n = int(L.ref_driver_nrays())
    ref.driver_finalize(variant=variant)
    return times, n
