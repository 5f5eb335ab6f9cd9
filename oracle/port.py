"""oracle/port.py -- TEST INFRASTRUCTURE ONLY.

ctypes client of ``oracle/liboracle_port.so`` = the CPU restatement of the hot path in ``oracle/port/calclens_port.c``.
Same interface as ``oracle/ref.py`` so tests can use either.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_port.so")
NAME = "port (oracle/port/calclens_port.c)"

RAY_DTYPE = np.dtype([("nest", "<i8"), ("n", "<f8", 3), ("beta", "<f8", 3), ("alpha", "<f8", 2),
                      ("A", "<f8", 4), ("Aprev", "<f8", 4), ("U", "<f8", 4), ("phi", "<f8")], align=False)
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        vp = C.c_void_p
        L.port_map2alm.restype = None; L.port_map2alm.argtypes = [C.c_long, C.c_long, vp, vp, vp, vp]
        L.port_poisson_filter.restype = None; L.port_poisson_filter.argtypes = [C.c_long, vp, vp]
        L.port_alm2allmaps.restype = None; L.port_alm2allmaps.argtypes = [C.c_long, C.c_long, vp, vp, vp]
        L.port_rayprop.restype = None; L.port_rayprop.argtypes = [vp, C.c_long, C.c_double, C.c_double, C.c_double]
        L.port_shearinterp.restype = None; L.port_shearinterp.argtypes = [C.c_long, vp, vp, C.c_long]
        L.port_init_rays.restype = None; L.port_init_rays.argtypes = [vp, C.c_long, C.c_long, C.c_long, C.c_double]
        L.port_plmgen.restype = C.c_long; L.port_plmgen.argtypes = [C.c_long, C.c_double, C.c_double, C.c_long, vp]
        for f in ("port_ring2nest", "port_nest2ring", "port_nest2peano"):
            getattr(L, f).restype = C.c_long; getattr(L, f).argtypes = [C.c_long, C.c_long]
        L.port_ang2nest.restype = C.c_long; L.port_ang2nest.argtypes = [C.c_double, C.c_double, C.c_long]
        L.port_nest2vec.restype = None; L.port_nest2vec.argtypes = [C.c_long, vp, C.c_long]
        L.port_get_interpol.restype = None; L.port_get_interpol.argtypes = [C.c_double, C.c_double, vp, vp, C.c_long]
        L.port_rayprop_born.restype = None; L.port_rayprop_born.argtypes = [vp, C.c_long, C.c_double, C.c_double, C.c_double]
        L.port_ray_output.restype = None; L.port_ray_output.argtypes = [vp, C.c_long, C.c_long]
        L.port_deposit_ngp.restype = None; L.port_deposit_ngp.argtypes = [vp, vp, C.c_long, C.c_long, vp]
        L.port_sizeof_ray.restype = C.c_long
        assert L.port_sizeof_ray() == 176
        _lib = L
    return _lib


def nlm(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def map2alm(order, lmax, ringmap, ring_weights=None):
    m = np.ascontiguousarray(ringmap, dtype=np.float32)
    are = np.zeros(nlm(lmax)); aim = np.zeros(nlm(lmax))
    w = None if ring_weights is None else np.ascontiguousarray(ring_weights, dtype=np.float64)
    lib().port_map2alm(order, lmax, None if w is None else w.ctypes.data, m.ctypes.data, are.ctypes.data, aim.ctypes.data)
    return are, aim


def poisson_filter(lmax, are, aim):
    are = np.array(are, dtype=np.float64); aim = np.array(aim, dtype=np.float64)
    lib().port_poisson_filter(lmax, are.ctypes.data, aim.ctypes.data)
    return are, aim


def alm2allmaps(order, lmax, are, aim):
    are = np.ascontiguousarray(are, dtype=np.float64); aim = np.ascontiguousarray(aim, dtype=np.float64)
    maps = np.zeros((6, 12 << (2 * order)), dtype=np.float32)
    lib().port_alm2allmaps(order, lmax, are.ctypes.data, aim.ctypes.data, maps.ctypes.data)
    return maps


def rayprop(rays, wp, wpm1, wpm2):
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    lib().port_rayprop(rays.ctypes.data, rays.size, wp, wpm1, wpm2)


def rayprop_born(rays, wp, wpm1, wpm2):
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    lib().port_rayprop_born(rays.ctypes.data, rays.size, wp, wpm1, wpm2)


def shearinterp(poisson_order, bundle_order, maps, rays):
    maps = np.ascontiguousarray(maps, dtype=np.float32)
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    lib().port_shearinterp(poisson_order, maps.ctypes.data, rays.ctypes.data, rays.size)


def init_rays(ray_order, binL_2, first=0, n=None):
    npix = 12 << (2 * ray_order)
    n = npix - first if n is None else n
    rays = np.zeros(n, dtype=RAY_DTYPE)
    lib().port_init_rays(rays.ctypes.data, first, n, ray_order, binL_2)
    return rays


def ray_output(rays, ray_order):
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    lib().port_ray_output(rays.ctypes.data, rays.size, ray_order)


def deposit_ngp(pos, mass, order):
    pos = np.ascontiguousarray(pos, dtype=np.float32); mass = np.ascontiguousarray(mass, dtype=np.float32)
    out = np.zeros(12 << (2 * order), dtype=np.float32)
    lib().port_deposit_ngp(pos.ctypes.data, mass.ctypes.data, mass.size, order, out.ctypes.data)
    return out


def plmgen(lmax, cth, sth, m):
    vec = np.zeros(lmax + 1)
    firstl = lib().port_plmgen(lmax, cth, sth, m, vec.ctypes.data)
    return firstl, vec


def ring2nest(p, order): return lib().port_ring2nest(int(p), order)
def nest2ring(p, order): return lib().port_nest2ring(int(p), order)
def nest2peano(p, order): return lib().port_nest2peano(int(p), order)
def ang2nest(t, p, order): return lib().port_ang2nest(float(t), float(p), order)


def get_interpol(theta, phi, order):
    pix = (C.c_long * 4)(); wgt = (C.c_double * 4)()
    lib().port_get_interpol(float(theta), float(phi), pix, wgt, order)
    return list(pix), list(wgt)
