"""Per-plane driver: the work of ``do_healpix_sht_poisson_solve`` (shtpoissonsolve.c:38-708, map-input path) followed by
the plane's ``rayprop_sphere`` calls (raytrace.c:256-269), on one GPU or sharded over the ranks of a
``torch.distributed`` process group (one process per GPU).

Sharding follows the reference: the map side of the SHT is split by ring pairs, the alm side
by m, joined by one transpose per direction (map2alm_transpose_mpi.c:339-381, alm2allmaps_transpose_mpi.c:656-724);
rays are split into contiguous NEST ranges, i.e. compact sky domains (cf. loadbalance.c:151-181).  Instead of the
reference's ring->domain shuffle with halo cells (map_shuffle.c) every rank's rings are stored into the map buffers of
the ranks whose ray domain + halo (``halo_deg``, default 1 degree) can reach them; with a halo a rank's ``maps`` are
therefore valid on its own rings and its halo region only (``halo_deg=0`` broadcasts full maps), and the ray kernel
raises an error if a ray's stencil leaves that region.

The driver itself lives below the C ABI (``clb_solver_*``, csrc/solver.cu): this class is a thin caller of it, so a C
host (CALCLENS through shim/calclens_b200_shim.c) reaches exactly the same code.  Two exchange back ends:
* fused (default on one NVLink/NVSwitch node, the C solver): the buffers of all ranks are mapped into every process
  (CUDA IPC) and the producing kernels -- Legendre-synthesis epilogue, map broadcast -- store straight into the
  consumer's buffers (the Legendre analysis pulls g from the ring owners); stages are ordered by a device-side barrier
  over peer memory, no communication library on the per-plane path.
* NCCL (``fused=False``; Python, over the ``_dev`` stage API): one all-to-all-v per transpose and an all-reduce of the
  maps (disjoint ring sets); the comparison arm, and the fallback when peer mapping is unavailable.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib, sht
from .rays import MODE_INTERP, MODE_PROP, MODE_ZERO

CSOL = 299792.458          # raytrace.h:110
RHO_CRIT = 2.77519737e11   # raytrace.h:109
MASS_SCALE = 1e10          # the reference's internal mass unit (shtpoissonsolve.c:426,468)


class Cosmology:
    """a(w) for flat LCDM exactly as the reference tabulates it (cosmocalc.c:14-92): 20000-point table of
    w(a) = 2997.92458 * int_a^1 da'/sqrt(a' Om + a'^4 (1-Om)), linear interpolation."""

    N, AMIN, AMAX = 20000, 0.01, 1.0

    def __init__(self, omega_m):
        self.omega_m = float(omega_m)
        i = np.arange(self.N - 1, dtype=np.float64)
        a = (self.AMAX - self.AMIN) / (self.N - 1.0) * i + self.AMIN
        x, w = np.polynomial.legendre.leggauss(96)
        # Gauss-Legendre on [a, 1] (the integrand is smooth for a >= 0.01); the reference uses gsl_integration_qag
        # with relerr 1e-8 (cosmocalc.c:42)
        half = 0.5 * (1.0 - a)[:, None]
        mid = 0.5 * (1.0 + a)[:, None]
        t = mid + half * x[None, :]
        f = 1.0 / np.sqrt(t * self.omega_m + t ** 4 * (1.0 - self.omega_m))
        self.aexpn = np.concatenate([a, [1.0]])
        self.comv = np.concatenate([(half[:, 0] * (f * w[None, :]).sum(axis=1)) * 2997.92458, [0.0]])

    def acomvdist(self, dist):
        """cosmocalc.c:57-92"""
        c = self.comv
        if dist < c[-1]:
            return self.AMAX
        if dist > c[0]:
            return self.AMIN
        # the reference's loop (cosmocalc.c:80-84) leaves i at the largest index with comv[i] > dist
        idx = np.nonzero(c > dist)[0]
        i = int(idx[-1]) if idx.size else 0
        i = max(i, 1)
        w = (dist - c[i - 1]) / (c[i] - c[i - 1])
        return (1.0 - w) * self.aexpn[i - 1] + w * self.aexpn[i]


def plane_params(plane, num_planes, max_comv_distance, omega_m, cosmo=None, pointmass=False, nobackdens=False):
    """set_plane_params (raytrace.c:384-423): returns dict(wpm1, wp, wpp1, densfact, backdens, z).
    Note the reference's naming at the call site raytrace.c:262: rayprop_sphere(planeRadPlus1, planeRad, planeRadMinus1)."""
    cosmo = cosmo or Cosmology(omega_m)
    binL = max_comv_distance / float(num_planes)
    wm1 = 0.0 if plane - 1 < 0 else (plane - 1.0) * binL + binL / 2.0
    w = plane * binL + binL / 2.0
    wp1 = max_comv_distance if plane + 1 == num_planes else (plane + 1.0) * binL + binL / 2.0
    if pointmass:
        radialvolume = w * w * binL
    else:
        radialvolume = (math.pow(w + binL / 2.0, 3.0) - math.pow(w - binL / 2.0, 3.0)) / 3.0
    zw = 1.0 / cosmo.acomvdist(w) - 1.0
    densfact = 3.0 * 100.0 * 100.0 / CSOL / CSOL * omega_m * w * (1.0 + zw) * binL / (radialvolume * RHO_CRIT * omega_m)
    backdens = 0.0 if nobackdens else 3.0 * 100.0 * 100.0 / CSOL / CSOL * omega_m * w * (1.0 + zw) * binL
    return dict(wpm1=wm1, wp=w, wpp1=wp1, densfact=densfact, backdens=backdens, z=zw, binL=binL)


def density_scalings(order, part_mass, densfact, backdens):
    """The three float factors the reference applies to a raw count map (shtpoissonsolve.c:426,468,478)."""
    area = 4.0 * math.pi / (12 << (2 * order))
    return (np.float32(part_mass / MASS_SCALE), np.float32(densfact / area * MASS_SCALE), np.float32(backdens))


class _SolverPlan(sht.HEALPixSHTPlan):
    """View of the plan owned by a C solver (not destroyed from Python)."""

    def __init__(self, lib, handle, order, lmax, nranks, rank, device):
        self.lib = lib
        self.order, self.lmax = int(order), int(lmax)
        self.nside = sht.order2nside(order)
        self.npix = sht.order2npix(order)
        self.nranks, self.rank = int(nranks), int(rank)
        self.device = device
        self.ring_weights = None
        self._h = handle
        self._query()

    def destroy(self):
        self._h = None


_ALLGATHER_T = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p)


class LensPlaneSolver:
    """One rank's share of the per-plane hot path.  ``dist_group`` = None for a single GPU, otherwise an initialised
    torch.distributed process group with one rank per GPU (used for bootstrap only on the fused path)."""

    def __init__(self, sht_order, lmax=None, ray_order=None, ring_weights=None, dist_group=None, device=None, fused=True,
                 halo_deg=1.0, allgather=None, nranks=None, rank=None, rp_owner=None, m_owner=None):
        import torch.distributed as dist
        self.dist = dist if dist_group is not None else None
        self.group = dist_group
        if dist_group is not None:
            self.nranks, self.rank = dist.get_world_size(dist_group), dist.get_rank(dist_group)
        else:
            self.nranks, self.rank = int(nranks or 1), int(rank or 0)   # (emulated ranks: allgather= a Python callable)
        self.lib = _lib.load()
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib.clb_set_device(self.device.index)
        self.order = int(sht_order)
        self.lmax = int(sht.order2lmax(self.order) if lmax is None else lmax)
        self.npix = sht.order2npix(self.order)
        self.ray_order = self.order if ray_order is None else int(ray_order)
        self.halo_deg = float(halo_deg)
        self.coarse_order = 5
        self.rays = None
        self.nrays = 0
        self.first_nest = 0
        self.fused = False
        self._need = None
        self._cs = None
        self._next = None
        self._timing = False
        w = None if ring_weights is None else np.ascontiguousarray(ring_weights, dtype=np.float64)
        ro = None if rp_owner is None else np.ascontiguousarray(rp_owner, dtype=np.int32)
        mo = None if m_owner is None else np.ascontiguousarray(m_owner, dtype=np.int32)
        if self.nranks == 1 or fused:
            gather = allgather
            if gather is None and self.nranks > 1:
                def gather(data):
                    out = [None] * self.nranks
                    self.dist.all_gather_object(out, data, group=self.group)
                    return out

            def _cb(send, recv, nbytes, _ctx):
                parts = gather(C.string_at(send, nbytes))
                C.memmove(recv, b"".join(parts), nbytes * self.nranks)
            self._cb = _ALLGATHER_T(_cb)   # keep the trampoline alive as long as the solver
            self._cs = self.lib.clb_solver_create(self.order, self.lmax, self.ray_order, None if w is None else w.ctypes.data,
                                                  self.nranks, self.rank, None if ro is None else ro.ctypes.data,
                                                  None if mo is None else mo.ctypes.data,
                                                  self._cb if self.nranks > 1 else _ALLGATHER_T(), None, self.halo_deg)
        if self._cs:
            q = lambda k: self.lib.clb_solver_query(self._cs, k)
            ptr = lambda k: self.lib.clb_solver_ptr(self._cs, k)
            self.fused = bool(q(2))
            self.plan = _SolverPlan(self.lib, ptr(4), self.order, self.lmax, self.nranks, self.rank, self.device)
            self.maps = self._dev_view(ptr(0), (6, self.npix), torch.float32)
            self.shells = int(q(8))
            self.maps2 = self._dev_view(ptr(9), (6, self.npix), torch.float32) if self.shells >= 2 else None
            self.alm_re = self._dev_view(ptr(2), (max(self.plan.Nlm, 1),), torch.float64)
            self.alm_im = self._dev_view(ptr(3), (max(self.plan.Nlm, 1),), torch.float64)
            self.summary = self._dev_view(ptr(8), (6,), torch.float64)
            if q(4):
                self._need = self._dev_view(ptr(5), (12 << (2 * self.coarse_order),), torch.uint8)
                self.need_fraction = q(5) * 1e-6
            self.host_barriers = bool(q(6))
            return
        # ---- NCCL exchange over the _dev stage API (fused=False, or peer mapping unavailable) ----
        self.plan = sht.HEALPixSHTPlan(sht_order, self.lmax, w, self.nranks, self.rank, rp_owner=ro, m_owner=mo, device=device)
        p = self.plan
        f64 = dict(dtype=torch.float64, device=self.device)
        self.g_send = torch.empty(2 * max(p.g_send_total, 1), **f64)
        self.b_send = torch.empty(2 * max(p.b_send_total, 1), **f64)
        self.g_recv = torch.empty(2 * max(p.g_recv_total, 1), **f64)
        self.b_recv = torch.empty(2 * max(p.b_recv_total, 1), **f64)
        self.alm_re = torch.empty(max(p.Nlm, 1), **f64)
        self.alm_im = torch.empty(max(p.Nlm, 1), **f64)
        self.summary = torch.zeros(6, **f64)
        self._dens = torch.zeros(self.npix, dtype=torch.float32, device=self.device)
        self.maps = torch.zeros((6, self.npix), dtype=torch.float32, device=self.device)

    def _dev_view(self, ptr, shape, dtype):
        """torch tensor over device memory owned by the library (CUDA array interface, zero copy)."""
        class _Holder:
            pass
        h = _Holder()
        ts = {torch.float64: "<f8", torch.float32: "<f4", torch.uint8: "|u1"}[dtype]
        h.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": ts, "data": (int(ptr), False), "version": 2,
                                      "strides": None}
        return torch.as_tensor(h, device=self.device)

    def close(self):
        """Release the solver (collective on multi-rank solvers: every rank must call it)."""
        if self._cs:
            torch.cuda.synchronize()
            self.maps = self.maps2 = self.alm_re = self.alm_im = self.summary = self._need = self.rays = None
            self.plan.destroy()
            self.lib.clb_solver_destroy(self._cs)
            self._cs = None
            self.fused = False

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    # ---- rays ----
    def init_rays(self, binL_2):
        """alloc_rays + init_rays (raytrace_utils.c:265-347) for this rank's contiguous NEST range."""
        if self._cs:
            self.nrays = self.lib.clb_solver_init_rays(self._cs, float(binL_2), self._stream())
            self.first_nest = self.lib.clb_solver_query(self._cs, 1)
            self.rays = self._dev_view(self.lib.clb_solver_ptr(self._cs, 1), (max(self.nrays, 1) * 176,), torch.uint8)
            return self.nrays
        nray_tot = 12 << (2 * self.ray_order)
        lo = (nray_tot * self.rank) // self.nranks
        hi = (nray_tot * (self.rank + 1)) // self.nranks
        self.first_nest, self.nrays = lo, hi - lo
        self.rays = torch.empty(max(self.nrays, 1) * 176, dtype=torch.uint8, device=self.device)
        self.lib.clb_ray_init_dev(self.rays.data_ptr(), self.nrays, lo, self.ray_order, float(binL_2), self._stream())
        return self.nrays

    def set_timing(self, on=True):
        """Per-stage CUDA events inside the C solver (bench.py)."""
        self._timing = bool(on)
        if self._cs:
            self.lib.clb_solver_set_timing(self._cs, 1 if on else 0)

    def stage_ms(self):
        out = (C.c_double * 9)()
        self.lib.clb_solver_stage_ms(self._cs, out)
        return list(out)

    def _all_to_all(self, send, recv, send_counts, recv_counts):
        self.dist.all_to_all_single(recv[:2 * sum(recv_counts)], send[:2 * sum(send_counts)],
                                    output_split_sizes=[2 * c for c in recv_counts],
                                    input_split_sizes=[2 * c for c in send_counts], group=self.group)
        return recv

    # ---- the SHT Poisson solve: counts map -> six derivative maps in self.maps ----
    def load_density(self, counts_map, premul, densmul, backdens):
        """Scale this rank's rings of a full-sky count map (device tensor or pinned host tensor, RING float32) into the
        density buffer: shtpoissonsolve.c:342-502 for the raw-map input path."""
        assert counts_map.dtype == torch.float32 and counts_map.numel() == self.npix and counts_map.is_contiguous()
        if self._cs:
            self.lib.clb_solver_load_density(self._cs, counts_map.data_ptr(), float(premul), float(densmul), float(backdens), self._stream())
            return None
        assert counts_map.is_cuda or counts_map.is_pinned(), "host maps must be pinned (they are read by the GPU directly)"
        self.lib.clb_load_density_dev(self.plan._h, counts_map.data_ptr(), self._dens.data_ptr(), float(premul), float(densmul),
                                      float(backdens), self._stream())
        return self._dens

    def solve(self, density=None):
        """density map (this rank's rings valid; default: the buffer load_density filled) -> six derivative maps."""
        if self._cs:
            self.lib.clb_solver_solve(self._cs, None if density is None else density.data_ptr(), self._stream())
            return self.maps
        p = self.plan
        dens = self._dens if density is None else density
        p.ring_analysis(dens, self.g_send)
        g = self._all_to_all(self.g_send, self.g_recv, p.counts[0], p.counts[1])
        p.legendre_analysis(g, self.alm_re, self.alm_im, poisson_filter=True)
        p.legendre_synthesis(self.alm_re, self.alm_im, self.b_send)
        b = self._all_to_all(self.b_send, self.b_recv, p.counts[2], p.counts[3])
        self.maps.zero_()
        p.ring_synthesis(b, self.maps)
        self.dist.all_reduce(self.maps, group=self.group)   # disjoint ring sets: x + 0 is exact
        return self.maps

    def alm2allmaps(self, alm_re, alm_im):
        """alm2allmaps_mpi over the ranks of the group: local alm (this rank's m) -> the six maps (fused exchange with a
        halo: every rank holds its own rings plus the part of the sky its rays can reach; otherwise full maps)."""
        if self._cs:
            self.lib.clb_solver_alm2allmaps(self._cs, alm_re.data_ptr(), alm_im.data_ptr(), self._stream())
            return self.maps
        p = self.plan
        p.legendre_synthesis(alm_re, alm_im, self.b_send)
        b = self._all_to_all(self.b_send, self.b_recv, p.counts[2], p.counts[3])
        self.maps.zero_()
        p.ring_synthesis(b, self.maps)
        self.dist.all_reduce(self.maps, group=self.group)
        return self.maps

    def ray_update(self, wpp1, wp, wpm1, with_summary=False):
        """zero + interpolate + propagate: rayprop_sphere(planeRadPlus1, planeRad, planeRadMinus1) as called at
        raytrace.c:262, preceded by the reset of raytrace.c:213-230 and the interpolation of shtpoissonsolve.c:666-702.
        with_summary: also accumulate the six plane sums into self.summary (same kernel, no second pass)."""
        if self._cs:
            self.lib.clb_solver_ray_update(self._cs, float(wpp1), float(wp), float(wpm1), MODE_ZERO | MODE_INTERP | MODE_PROP,
                                           1 if with_summary else 0, self._stream())
            return
        ptrs = (C.c_void_p * 6)(*[self.maps[k].data_ptr() for k in range(6)])
        self.lib.clb_ray_step_ex_dev(self.rays.data_ptr(), self.nrays, ptrs, self.order, float(wpp1), float(wp), float(wpm1),
                                     MODE_ZERO | MODE_INTERP | MODE_PROP, None, 0, self.rank, None,
                                     self.summary.data_ptr() if with_summary else None, self._stream())

    def check_halo(self):
        """Raise if a ray left the part of the sky this rank receives (the reference aborts on a missing map cell,
        shtpoissonsolve.c:683-689); synchronises."""
        if self._cs:
            self._raise(self.lib.clb_solver_check(self._cs, self._stream()))

    def _raise(self, err):
        if err & 2:
            raise RuntimeError("calclens_b200: a rank never reached a barrier of the fused exchange")
        if err & 1:
            raise RuntimeError("calclens_b200: a ray left its domain + halo (%.2f deg); raise halo_deg" % self.halo_deg)

    def step(self, counts_map, premul, densmul, backdens, wpp1, wp, wpm1, read_summary=True, prefetch=None, pair=None):
        """One lens plane.  ``counts_map``: RING float32 full-sky map, a device tensor or a (pinned) host tensor.
        ``prefetch`` = (next_counts_map, premul, densmul, backdens), or a list of up to two of them, starts the load of the
        planes to come behind this plane's kernels; a later step() given that same map picks the staged density up
        instead of loading again.
        ``pair`` = (partner_counts_map, premul, densmul, backdens): the plane AFTER this one.  Both go through one pass of
        each Legendre kernel (two shells per pass); the next step(), given the partner map, only updates the rays.
        Returns the six ray sums over all ranks (read_summary) or None."""
        assert counts_map.dtype == torch.float32 and counts_map.numel() == self.npix and counts_map.is_contiguous()
        if self._cs:
            if pair is not None:
                nm, a, b, c = pair
                self._pair = nm
                self.lib.clb_solver_set_pair(self._cs, nm.data_ptr(), float(a), float(b), float(c))
            if prefetch is not None:
                if isinstance(prefetch, tuple):
                    prefetch = [prefetch]
                self._next = [x[0] for x in prefetch]    # keep the tensors alive until their loads have run
                for nm, a, b, c in prefetch:
                    self.lib.clb_solver_set_next(self._cs, nm.data_ptr(), float(a), float(b), float(c))
            out = (C.c_double * 6)() if read_summary else None
            err = self.lib.clb_solver_step(self._cs, counts_map.data_ptr(), float(premul), float(densmul), float(backdens),
                                           float(wpp1), float(wp), float(wpm1), out, self._stream())
            if not read_summary:
                return None
            self._raise(err)
            res = np.array(list(out))
            if self.nranks > 1 and self.dist is not None:
                t = torch.from_numpy(res).to(self.device)
                self.dist.all_reduce(t, group=self.group)
                res = t.cpu().numpy()
            return res
        assert pair is None, "two shells per pass need the fused C solver"
        self.load_density(counts_map, premul, densmul, backdens)
        self.solve()
        self.ray_update(wpp1, wp, wpm1, with_summary=read_summary)
        if not read_summary:
            return None
        self.dist.all_reduce(self.summary, group=self.group)
        return self.summary.cpu().numpy()

    def rays_host(self):
        from .rays import rays_from_device
        return rays_from_device(self.rays[:self.nrays * 176])

    def sync_rays(self):
        """(kept for callers of the round-1 API: the ray kernel runs on the caller's stream, nothing to join)"""
