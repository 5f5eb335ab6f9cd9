"""calclens_b200: sm_100a CUDA implementation of the CALCLENS SHTONLY lens-plane hot path
(spherical-harmonic Poisson solve + per-ray deflection / shear update) behind the reference's function boundary.

The compute lives in ``libcalclens_b200.so`` (C ABI: ``include/calclens_b200.h``); this package is the host-side
mirror of the reference interface for that path.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .sht import (HEALPixSHTPlan, alm2allmaps_mpi, default_owners, lm2index, map2alm_mpi, num_lms, order2lmax,  # noqa: F401
                  order2nside, order2npix, read_ring_weights)
from .rays import RAY_DTYPE, rayprop_sphere, shearinterp_rays  # noqa: F401

__all__ = ["HEALPixSHTPlan", "map2alm_mpi", "alm2allmaps_mpi", "rayprop_sphere", "shearinterp_rays", "RAY_DTYPE",
           "read_ring_weights", "order2lmax", "order2nside", "order2npix", "num_lms", "lm2index", "default_owners"]
