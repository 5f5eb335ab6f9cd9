"""Host-side mirror of the reference's ray interface (raytrace.h:284-293, :431) over the CUDA library."""
import ctypes as C

import numpy as np
import torch

from . import _lib

# HEALPixRay, raytrace.h:284-293 (176 bytes)
RAY_DTYPE = np.dtype([("nest", "<i8"), ("n", "<f8", 3), ("beta", "<f8", 3), ("alpha", "<f8", 2),
                      ("A", "<f8", 4), ("Aprev", "<f8", 4), ("U", "<f8", 4), ("phi", "<f8")], align=False)
assert RAY_DTYPE.itemsize == 176

MODE_ZERO, MODE_INTERP, MODE_PROP, MODE_BORN = 1, 2, 4, 8


def rayprop_sphere(wp, wpm1, wpm2, rays, born=False):
    """rayprop_sphere(wp, wpm1, wpm2, bundleCellInd) (rayprop.c:18) over a host array of HEALPixRay, in place.
    The reference is called once per bundle cell (raytrace.c:256-269); here the caller passes the cell's rays
    (or all rays at once).  born=True: the reference's -DBORNAPPRX build (rayprop.c:40-62)."""
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    _lib.load().clb_ray_step(rays.ctypes.data, rays.size, None, 0, wp, wpm1, wpm2, MODE_PROP | (MODE_BORN if born else 0))


def shearinterp_rays(maps, map_order, rays, zero_first=False):
    """The per-ray loop of do_healpix_sht_poisson_solve (shtpoissonsolve.c:666-702): interpolate the six RING maps at
    every ray and accumulate phi, alpha, U; host arrays, in place."""
    assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
    m = np.ascontiguousarray(maps, dtype=np.float32)
    assert m.shape == (6, 12 << (2 * map_order))
    _lib.load().clb_ray_step(rays.ctypes.data, rays.size, m.ctypes.data, map_order, 0.0, 0.0, 0.0,
                             MODE_INTERP | (MODE_ZERO if zero_first else 0))


def ray_step_dev(rays_dev, nrays, maps_dev, map_order, wp, wpm1, wpm2, mode):
    """Device-resident variant: rays_dev is a uint8/int64 torch tensor holding nrays 176-byte records."""
    L = _lib.load()
    ptrs = None
    if maps_dev is not None:
        ptrs = (C.c_void_p * 6)(*[maps_dev[k].data_ptr() for k in range(6)])
    return L.clb_ray_step_dev(rays_dev.data_ptr(), nrays, ptrs, map_order, wp, wpm1, wpm2, mode,
                              torch.cuda.current_stream().cuda_stream)


def rays_to_device(rays, device=None):
    t = torch.from_numpy(rays.view(np.uint8).reshape(-1))
    return t.to(device or torch.device("cuda", torch.cuda.current_device()))


def rays_from_device(t):
    return t.cpu().numpy().view(RAY_DTYPE)
