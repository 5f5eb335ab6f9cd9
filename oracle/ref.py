"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes client of ``oracle/_ref/libcalclens_ref.so`` = the UNMODIFIED CALCLENS reference functions compiled from
/root/reference by ``oracle/Makefile`` (single-rank MPI stub, FP64-internal float FFT shim).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; the product package
``calclens_b200`` never does.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libcalclens_ref.so")

# HEALPixRay, raytrace.h:284-293 (176 bytes)
RAY_DTYPE = np.dtype([("nest", "<i8"), ("n", "<f8", 3), ("beta", "<f8", 3), ("alpha", "<f8", 2),
                      ("A", "<f8", 4), ("Aprev", "<f8", 4), ("U", "<f8", 4), ("phi", "<f8")], align=False)
assert RAY_DTYPE.itemsize == 176

_VARIANTS = {"ref": "libcalclens_ref.so",            # parity build: -O2 -ffp-contract=off
             "fast": "libcalclens_ref_fast.so",      # timing only: the reference's own -O3 -ffast-math -funroll-loops
             "shim": "libcalclens_ref_shim.so"}      # same harness linked against shim/calclens_b200_shim.c (GPU)
_libs = {}


def path(variant="ref"):
    return os.path.join(_HERE, "_ref", _VARIANTS[variant])


def available(variant="ref"):
    return os.path.exists(path(variant))


def lib(variant="ref"):
    if variant not in _libs:
        if not available(variant):
            raise RuntimeError("oracle/_ref/%s missing: run `make -C oracle %s` where /root/reference exists" % (
                _VARIANTS[variant], {"ref": "ref", "fast": "fast", "shim": "shim"}[variant]))
        L = C.CDLL(path(variant))
        dp, fp, lp, vp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_long), C.c_void_p
        L.ref_nmapvec.restype = C.c_long; L.ref_nmapvec.argtypes = [C.c_long]
        L.ref_nlm.restype = C.c_long; L.ref_nlm.argtypes = [C.c_long]
        L.ref_map2alm.restype = None; L.ref_map2alm.argtypes = [C.c_long, C.c_long, vp, vp, vp, vp]
        L.ref_poisson_filter.restype = None; L.ref_poisson_filter.argtypes = [C.c_long, vp, vp]
        L.ref_alm2allmaps.restype = None; L.ref_alm2allmaps.argtypes = [C.c_long, C.c_long, vp, vp, vp]
        L.ref_rayprop.restype = None; L.ref_rayprop.argtypes = [vp, C.c_long, C.c_double, C.c_double, C.c_double]
        L.ref_shearinterp.restype = C.c_long; L.ref_shearinterp.argtypes = [C.c_long, C.c_long, vp, vp, C.c_long]
        L.ref_read_ring_weights.restype = C.c_long; L.ref_read_ring_weights.argtypes = [C.c_char_p, C.c_long, vp]
        L.ref_plmgen.restype = C.c_long; L.ref_plmgen.argtypes = [C.c_long, C.c_double, C.c_double, C.c_long, vp]
        L.ref_sizeof_ray.restype = C.c_long
        L.ref_init_rays.restype = None; L.ref_init_rays.argtypes = [vp, C.c_long, C.c_long, C.c_long, C.c_double]
        L.ref_rayprop_born.restype = None; L.ref_rayprop_born.argtypes = [vp, C.c_long, C.c_double, C.c_double, C.c_double]
        L.ref_ray_output.restype = None; L.ref_ray_output.argtypes = [vp, C.c_long, C.c_long]
        L.ref_deposit_ngp.restype = None; L.ref_deposit_ngp.argtypes = [vp, vp, C.c_long, C.c_long, vp]
        L.ref_is_shim.restype = C.c_int
        L.ref_driver_init.restype = None
        L.ref_driver_init.argtypes = [C.c_long, C.c_long, C.c_long, C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_long, C.c_double, C.c_char_p]
        L.ref_driver_plane.restype = None; L.ref_driver_plane.argtypes = [C.c_long] + [C.c_double] * 5
        L.ref_driver_nrays.restype = C.c_long
        L.ref_driver_get_rays.restype = None; L.ref_driver_get_rays.argtypes = [vp]
        L.ref_driver_finalize.restype = None
        L.ref_mpi_rank.restype = C.c_int; L.ref_mpi_size.restype = C.c_int
        L.ref_mpi_plan_info.restype = None; L.ref_mpi_plan_info.argtypes = [C.c_long, C.c_long, vp]
        L.ref_mpi_map2alm.restype = None; L.ref_mpi_map2alm.argtypes = [C.c_long, C.c_long, vp, vp, vp, vp]
        L.ref_mpi_alm2allmaps.restype = None; L.ref_mpi_alm2allmaps.argtypes = [C.c_long, C.c_long, vp, vp, vp]
        L.ref_sample_map2alm.restype = None; L.ref_sample_map2alm.argtypes = [C.c_long, C.c_long, vp, vp, C.c_long, vp, vp, vp]
        L.ref_sample_alm2allmaps_rings.restype = None
        L.ref_sample_alm2allmaps_rings.argtypes = [C.c_long, C.c_long, vp, vp, C.c_long, vp, vp]
        assert L.ref_sizeof_ray() == 176
        # HEALPix helpers straight from healpix_utils.c
        L.ring2nest.restype = C.c_long; L.ring2nest.argtypes = [C.c_long, C.c_long]
        L.nest2ring.restype = C.c_long; L.nest2ring.argtypes = [C.c_long, C.c_long]
        L.ang2nest.restype = C.c_long; L.ang2nest.argtypes = [C.c_double, C.c_double, C.c_long]
        L.ang2ring.restype = C.c_long; L.ang2ring.argtypes = [C.c_double, C.c_double, C.c_long]
        L.nest2peano.restype = C.c_long; L.nest2peano.argtypes = [C.c_long, C.c_long]
        L.peano2nest.restype = C.c_long; L.peano2nest.argtypes = [C.c_long, C.c_long]
        L.nest2vec.restype = None; L.nest2vec.argtypes = [C.c_long, dp, C.c_long]
        L.vec2ang.restype = None; L.vec2ang.argtypes = [dp, dp, dp]
        L.get_interpol.restype = None; L.get_interpol.argtypes = [C.c_double, C.c_double, lp, dp, C.c_long]
        L.get_ring_info2.restype = None; L.get_ring_info2.argtypes = [C.c_long, lp, lp, dp, dp, lp, C.c_long]
        L.get_lmin_ylm.restype = C.c_long; L.get_lmin_ylm.argtypes = [C.c_long, C.c_double]
        _libs[variant] = L
    return _libs[variant]


def nlm(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def map2alm(order, lmax, ringmap, ring_weights=None, variant="ref"):
    """map2alm_mpi on a RING-ordered float32 map -> (alm_re, alm_im), m-major."""
    m = np.ascontiguousarray(ringmap, dtype=np.float32)
    assert m.size == 12 << (2 * order)
    are = np.zeros(nlm(lmax)); aim = np.zeros(nlm(lmax))
    w = None if ring_weights is None else np.ascontiguousarray(ring_weights, dtype=np.float64)
    lib(variant).ref_map2alm(order, lmax, None if w is None else w.ctypes.data, m.ctypes.data, are.ctypes.data, aim.ctypes.data)
    return are, aim


def poisson_filter(lmax, are, aim):
    are = np.array(are, dtype=np.float64); aim = np.array(aim, dtype=np.float64)
    lib().ref_poisson_filter(lmax, are.ctypes.data, aim.ctypes.data)
    return are, aim


def alm2allmaps(order, lmax, are, aim, variant="ref"):
    """alm2allmaps_mpi -> float32 array [6, Npix] RING-ordered (phi, gt, gp, gtt, gtp, gpp)."""
    are = np.ascontiguousarray(are, dtype=np.float64).copy(); aim = np.ascontiguousarray(aim, dtype=np.float64).copy()
    maps = np.zeros((6, 12 << (2 * order)), dtype=np.float32)
    lib(variant).ref_alm2allmaps(order, lmax, are.ctypes.data, aim.ctypes.data, maps.ctypes.data)
    return maps


def rayprop(rays, wp, wpm1, wpm2, variant="ref"):
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    lib(variant).ref_rayprop(rays.ctypes.data, rays.size, wp, wpm1, wpm2)


def rayprop_born(rays, wp, wpm1, wpm2):
    """rayprop_sphere of a -DBORNAPPRX build (rayprop.c:40-62)."""
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    lib().ref_rayprop_born(rays.ctypes.data, rays.size, wp, wpm1, wpm2)


def shearinterp(poisson_order, bundle_order, maps, rays, variant="ref"):
    maps = np.ascontiguousarray(maps, dtype=np.float32)
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    bad = lib(variant).ref_shearinterp(poisson_order, bundle_order, maps.ctypes.data, rays.ctypes.data, rays.size)
    if bad:
        raise RuntimeError("reference shearinterp_comp reported %d rays without map cells" % bad)


def read_ring_weights(path, order):
    out = np.zeros(2 << order)
    n = lib().ref_read_ring_weights(path.encode(), order, out.ctypes.data)
    if n != out.size:
        raise RuntimeError("ring weights not read from %s" % path)
    return out


def plmgen(lmax, cth, sth, m):
    vec = np.zeros(lmax + 1)
    firstl = lib().ref_plmgen(lmax, cth, sth, m, vec.ctypes.data)
    return firstl, vec


def init_rays(ray_order, binL_2, first=0, n=None):
    """Ray initialisation as raytrace_utils.c:302-347 (beta = pixel centre, n = beta*binL/2, A = Aprev = I)."""
    npix = 12 << (2 * ray_order)
    n = npix - first if n is None else n
    rays = np.zeros(n, dtype=RAY_DTYPE)
    lib().ref_init_rays(rays.ctypes.data, first, n, ray_order, binL_2)
    return rays


def ray_output(rays, ray_order):
    """write_rays' pre-output transform (rayio.c:300-312) in place: paratrans_ray_curr2obs + rot_ray_ang2radec."""
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    lib().ref_ray_output(rays.ctypes.data, rays.size, ray_order)


def deposit_ngp(pos, mass, order):
    """NGP deposit of shtpoissonsolve.c:128-150 -> RING float32 map (sequential float accumulation)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32); mass = np.ascontiguousarray(mass, dtype=np.float32)
    out = np.zeros(12 << (2 * order), dtype=np.float32)
    lib().ref_deposit_ngp(pos.ctypes.data, mass.ctypes.data, mass.size, order, out.ctypes.data)
    return out


# ---- sampled oracle (reference primitives + restated glue; see ref_harness.c) ----
def sample_map2alm(order, lmax, ringmap, mlist, ring_weights=None, variant="ref"):
    """alm rows of the selected m (list of arrays, index l - m)."""
    m = np.ascontiguousarray(ringmap, dtype=np.float32)
    ml = np.ascontiguousarray(sorted(int(x) for x in mlist), dtype=np.int64)
    n = int((lmax - ml + 1).sum())
    are = np.zeros(n); aim = np.zeros(n)
    w = None if ring_weights is None else np.ascontiguousarray(ring_weights, dtype=np.float64)
    lib(variant).ref_sample_map2alm(order, lmax, None if w is None else w.ctypes.data, m.ctypes.data, ml.size, ml.ctypes.data,
                                    are.ctypes.data, aim.ctypes.data)
    out, off = {}, 0
    for mm in ml:
        k = lmax - int(mm) + 1
        out[int(mm)] = (are[off:off + k].copy(), aim[off:off + k].copy()); off += k
    return out


def sample_alm2allmaps_rings(order, lmax, are, aim, rings, variant="ref"):
    """Six float maps on the selected ring pairs (north ring numbers 1..2Nside): dict ring -> (north [6, n], south [6, n] or None)."""
    nside = 1 << order
    rl = np.ascontiguousarray(sorted(int(r) for r in rings), dtype=np.int64)
    npx = [(4 * r if r < nside else 4 * nside) for r in rl]
    tot = sum(n * (1 if r == 2 * nside else 2) for n, r in zip(npx, rl))
    are = np.ascontiguousarray(are, dtype=np.float64); aim = np.ascontiguousarray(aim, dtype=np.float64)
    out = np.zeros((6, tot), dtype=np.float32)
    lib(variant).ref_sample_alm2allmaps_rings(order, lmax, are.ctypes.data, aim.ctypes.data, rl.size, rl.ctypes.data, out.ctypes.data)
    res, off = {}, 0
    for n, r in zip(npx, rl):
        north = out[:, off:off + n].copy(); off += n
        south = None
        if r != 2 * nside:
            south = out[:, off:off + n].copy(); off += n
        res[int(r)] = (north, south)
    return res


# ---- the reference's own plane loop on the raw-map input path (ref_harness.c: ref_driver_*) ----
def driver_init(bundle_order, ray_order, map_order, map_path, map_name, part_mass, max_comv_distance, num_planes, omega_m,
                ring_weight_path="", variant="ref"):
    lib(variant).ref_driver_init(bundle_order, ray_order, map_order, map_path.encode(), map_name.encode(), float(part_mass),
                                 float(max_comv_distance), int(num_planes), float(omega_m), ring_weight_path.encode())


def driver_plane(plane, wpm1, wp, wpp1, densfact, backdens, variant="ref"):
    lib(variant).ref_driver_plane(int(plane), float(wpm1), float(wp), float(wpp1), float(densfact), float(backdens))


def driver_rays(variant="ref"):
    L = lib(variant)
    rays = np.zeros(L.ref_driver_nrays(), dtype=RAY_DTYPE)
    L.ref_driver_get_rays(rays.ctypes.data)
    return rays


def driver_finalize(variant="ref"):
    lib(variant).ref_driver_finalize()


# ---- map2alm_mpi / alm2allmaps_mpi over the ranks of the shared-memory MPI stub (inside oracle/mpirun.py workers) ----
def mpi_plan_info(order, lmax, variant="ref"):
    info = (C.c_long * 5)()
    lib(variant).ref_mpi_plan_info(order, lmax, info)
    return dict(first_m=info[0], last_m=info[1], nlm=info[2], first_ring=info[3], last_ring=info[4])


def mpi_map2alm(order, lmax, ringmap, ring_weights=None, variant="ref"):
    info = mpi_plan_info(order, lmax, variant)
    m = np.ascontiguousarray(ringmap, dtype=np.float32)
    are = np.zeros(max(info["nlm"], 1)); aim = np.zeros(max(info["nlm"], 1))
    w = None if ring_weights is None else np.ascontiguousarray(ring_weights, dtype=np.float64)
    lib(variant).ref_mpi_map2alm(order, lmax, None if w is None else w.ctypes.data, m.ctypes.data, are.ctypes.data, aim.ctypes.data)
    return info, are[:info["nlm"]], aim[:info["nlm"]]


def mpi_alm2allmaps(order, lmax, are, aim, variant="ref"):
    """this rank's alm slice -> six RING maps with this rank's rings filled (zeros elsewhere)"""
    are = np.ascontiguousarray(are, dtype=np.float64).copy(); aim = np.ascontiguousarray(aim, dtype=np.float64).copy()
    maps = np.zeros((6, 12 << (2 * order)), dtype=np.float32)
    lib(variant).ref_mpi_alm2allmaps(order, lmax, are.ctypes.data, aim.ctypes.data, maps.ctypes.data)
    return maps


NAME = "reference (oracle/_ref)"


def ring2nest(p, order): return lib().ring2nest(int(p), order)
def nest2ring(p, order): return lib().nest2ring(int(p), order)
def nest2peano(p, order): return lib().nest2peano(int(p), order)
def ang2nest(t, p, order): return lib().ang2nest(float(t), float(p), order)


def get_interpol(theta, phi, order):
    pix = (C.c_long * 4)(); wgt = (C.c_double * 4)()
    lib().get_interpol(float(theta), float(phi), pix, wgt, order)
    return list(pix), list(wgt)
