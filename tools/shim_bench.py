"""Cost of the drop-in binding: the REFERENCE's own plane loop (ref_driver_*: domain decomposition, ray reset,
do_healpix_sht_poisson_solve, rayprop_sphere per bundle cell; oracle/ref_harness.c) linked against
shim/calclens_b200_shim.c, timed per plane on one rank / one GPU, with the rays round-tripping host <-> device in every
call (default binding, no host change) and device resident (CALCLENS_B200_RESIDENT=1).  Test/bench infrastructure.
    python tools/shim_bench.py [order=10] [ray_order=10] [bundle_order=5] [planes=4]
Prints one JSON line per mode."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    kw = dict(order=10, ray_order=10, bundle_order=5, planes=4)
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = int(v)
    from calclens_b200 import poisson
    from oracle import mpirun, ref
    from tests import mpi_workers
    if not ref.available("shim"):
        print(json.dumps({"unavailable": "oracle/_ref/libcalclens_ref_shim.so not built"}))
        return
    order, nplanes = kw["order"], kw["planes"]
    npix = 12 << (2 * order)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        rng = np.random.default_rng(5)
        for p in range(nplanes):
            (8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32).tofile(os.path.join(tmp, "lensmap.%d" % p))
        cosmo = poisson.Cosmology(0.27)
        max_dist = 30.0 * nplanes
        planes = []
        for p in range(nplanes):
            pp = poisson.plane_params(p, nplanes, max_dist, 0.27, cosmo)
            planes.append(dict(plane=p, wpm1=pp["wpm1"], wp=pp["wp"], wpp1=pp["wpp1"], densfact=pp["densfact"] * 1e3, backdens=pp["backdens"] * 1e3))
        cfg = dict(bundle_order=kw["bundle_order"], ray_order=kw["ray_order"], map_order=order, map_path=tmp, map_name="lensmap",
                   part_mass=3.0e10, max_comv_distance=max_dist, num_planes=nplanes, omega_m=0.27)
        for resident in (0, 1):
            times, nrays = mpirun.run(1, mpi_workers.timed_driver_planes, cfg, planes, "shim", timeout=1200,
                                      extra_env={"CALCLENS_B200_RESIDENT": str(resident)})[0]
            steady = times[1:] if len(times) > 1 else times     # the first plane pays plan creation (seeds, FFT tables)
            print(json.dumps({"binding": "reference plane loop through shim/calclens_b200_shim.c (one rank, one GPU)",
                              "rays": "device resident" if resident else "host <-> device round trip in every call",
                              "nside": 1 << order, "lmax": 3 * (1 << order) - 1, "ray_nside": 1 << kw["ray_order"], "nrays": nrays,
                              "seconds_per_plane": [round(t, 4) for t in times], "steady_ms_per_plane": 1e3 * float(np.mean(steady))}))


if __name__ == "__main__":
    main()
