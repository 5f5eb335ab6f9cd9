"""ctypes binding of libcalclens_b200.so (the C ABI in include/calclens_b200.h).

The library is the product; there is no Python or CPU fallback.  Import fails loudly if the shared object is
missing, and every compute call aborts inside the library if no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcalclens_b200.so")

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "calclens_b200: %s is missing. Build it with `python -m calclens_b200.build` (needs nvcc); "
            "there is no fallback implementation." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, lp, ip, dp = C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_int), C.POINTER(C.c_double)
    sig = {
        "clb_abi_version": (C.c_int, []),
        "clb_device_count": (C.c_int, []),
        "clb_set_device": (None, [C.c_int]),
        "clb_launch_count": (C.c_long, []),
        "clb_set_tuning": (None, [C.c_int, C.c_int]),
        "clb_sht_plan_create": (vp, [C.c_long, C.c_long, vp, C.c_int, C.c_int, vp, vp]),
        "clb_sht_plan_destroy": (None, [vp]),
        "clb_sht_plan_query": (C.c_long, [vp, C.c_int]),
        "clb_sht_plan_counts": (None, [vp, C.c_int, vp]),
        "clb_sht_plan_local_m": (None, [vp, vp]),
        "clb_sht_plan_local_ring_pairs": (None, [vp, vp]),
        "clb_peer_alloc": (vp, [C.c_long]),
        "clb_peer_free": (None, [vp]),
        "clb_peer_export": (None, [vp, vp]),
        "clb_peer_import": (vp, [vp]),
        "clb_peer_release": (None, [vp]),
        "clb_sht_plan_set_peers": (None, [vp, vp, vp]),
        "clb_sht_plan_set_peers_shells": (None, [vp, vp, vp, C.c_int]),
        "clb_maps_broadcast_dev": (C.c_int, [vp, vp, vp, vp, C.c_long, vp]),
        "clb_domain_masks": (None, [C.c_long, C.c_int, C.c_long, C.c_double, vp]),
        "clb_ray_step_ex_dev": (C.c_int, [vp, C.c_long, vp, C.c_long, C.c_double, C.c_double, C.c_double, C.c_int, vp, C.c_long, C.c_int, vp, vp, vp]),
        "clb_ring_analysis_dev": (C.c_int, [vp, vp, vp, vp]),
        "clb_legendre_analysis_dev": (C.c_int, [vp, vp, vp, vp, C.c_int, vp]),
        "clb_legendre_synthesis_dev": (C.c_int, [vp, vp, vp, vp, vp]),
        "clb_legendre_analysis_shells_dev": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp]),
        "clb_legendre_synthesis_shells_dev": (C.c_int, [vp, vp, vp, vp, C.c_int, vp]),
        "clb_ring_synthesis_dev": (C.c_int, [vp, vp, vp, vp]),
        "clb_scale_density_dev": (C.c_int, [vp, C.c_long, C.c_float, C.c_float, C.c_float, vp]),
        "clb_load_density_dev": (C.c_int, [vp, vp, vp, C.c_float, C.c_float, C.c_float, vp]),
        "clb_ray_step_dev": (C.c_int, [vp, C.c_long, vp, C.c_long, C.c_double, C.c_double, C.c_double, C.c_int, vp]),
        "clb_ray_init_dev": (C.c_int, [vp, C.c_long, C.c_long, C.c_long, C.c_double, vp]),
        "clb_ray_summary_dev": (C.c_int, [vp, C.c_long, vp, vp]),
        "clb_ray_output_dev": (C.c_int, [vp, vp, C.c_long, C.c_long, vp]),
        "clb_deposit_ngp_dev": (C.c_int, [vp, vp, C.c_long, C.c_long, vp, vp]),
        "clb_map2alm": (None, [vp, vp, vp, vp, C.c_int]),
        "clb_alm2allmaps": (None, [vp, vp, vp, vp]),
        "clb_map2alm_mapvec": (None, [vp, vp, vp, vp, vp, vp]),
        "clb_alm2allmaps_mapvec": (None, [vp, vp, vp, vp, vp, vp]),
        "clb_ray_step": (None, [vp, C.c_long, vp, C.c_long, C.c_double, C.c_double, C.c_double, C.c_int]),
        "clb_lens_plane": (None, [vp, vp, C.c_float, C.c_float, C.c_float, vp, C.c_long, C.c_double, C.c_double, C.c_double]),
        "clb_healpix_index_dev": (None, [C.c_int, C.c_long, C.c_long, vp, vp, vp, vp, vp]),
        "clb_healpix_interpol_dev": (None, [C.c_long, C.c_long, vp, vp, vp, vp]),
        "clb_ray_stencil_dev": (None, [C.c_long, C.c_long, vp, vp, vp, vp]),
        "clb_host_register": (C.c_int, [vp, C.c_long]),
        "clb_host_unregister": (None, [vp]),
        "clb_pool_release": (None, []),
        # persistent solver (csrc/solver.cu); the allgather callback is a CFUNCTYPE(None, void*, void*, long, void*)
        "clb_solver_create": (vp, [C.c_long, C.c_long, C.c_long, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_double]),
        "clb_solver_destroy": (None, [vp]),
        "clb_solver_query": (C.c_long, [vp, C.c_int]),
        "clb_solver_ptr": (vp, [vp, C.c_int]),
        "clb_solver_init_rays": (C.c_long, [vp, C.c_double, vp]),
        "clb_solver_set_rays": (None, [vp, vp, C.c_long, vp]),
        "clb_solver_get_rays": (None, [vp, vp, vp]),
        "clb_solver_ray_output": (None, [vp, vp, vp]),
        "clb_solver_step": (C.c_int, [vp, vp, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, C.c_double, vp, vp]),
        "clb_solver_set_next": (None, [vp, vp, C.c_float, C.c_float, C.c_float]),
        "clb_solver_set_pair": (None, [vp, vp, C.c_float, C.c_float, C.c_float]),
        "clb_solver_check": (C.c_int, [vp, vp]),
        "clb_solver_load_density": (None, [vp, vp, C.c_float, C.c_float, C.c_float, vp]),
        "clb_solver_solve": (None, [vp, vp, vp]),
        "clb_solver_alm2allmaps": (None, [vp, vp, vp, vp]),
        "clb_solver_ray_update": (None, [vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, vp]),
        "clb_solver_map2alm_mapvec": (None, [vp, vp, vp, vp, vp, vp, vp]),
        "clb_solver_alm2allmaps_mapvec": (None, [vp, vp, vp, vp, vp, vp, vp]),
        "clb_solver_set_timing": (None, [vp, C.c_int]),
        "clb_solver_stage_ms": (None, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    L._clb_signatures = sig
    _lib = L
    return L


def exported_symbols():
    """Names declared in include/calclens_b200.h (parsed), used by the CPU-side ABI test."""
    import re
    hdr = os.path.join(_HERE, "..", "include", "calclens_b200.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(clb_[a-z0-9_]+)\s*\(", txt)) - {"clb_allgather_fn"})
