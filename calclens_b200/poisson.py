"""Per-plane driver: the work of ``do_healpix_sht_poisson_solve`` (shtpoissonsolve.c:38-708, map-input path) followed by
the plane's ``rayprop_sphere`` calls (raytrace.c:256-269), on one GPU or sharded over the ranks of a
``torch.distributed`` process group (one process per GPU).

Sharding follows the reference: the map side of the SHT is split by ring pairs, the alm side
by m, joined by one transpose per direction (map2alm_transpose_mpi.c:339-381, alm2allmaps_transpose_mpi.c:656-724);
rays are split into contiguous NEST ranges, i.e. compact sky domains (cf. loadbalance.c:151-181).  Instead of the
reference's ring->domain shuffle with halo cells (map_shuffle.c) every rank receives the six full derivative maps, so
rays never miss a map cell however far they have been deflected.

Two exchange back ends:
* fused (default on one NVLink/NVSwitch node): the receive buffers of all ranks are mapped into every process (CUDA
  IPC) and the producing kernels -- ring-FFT epilogue, Legendre-synthesis epilogue, map broadcast -- store straight
  into the consumer's buffers; the only collectives left are stream-ordered barriers between producer and consumer.
* NCCL: one all-to-all-v per transpose and an all-reduce of the maps (disjoint ring sets); used when peer mapping is
  unavailable, and by the CPU (gloo) tests of the exchange layout.
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


class LensPlaneSolver:
    """One rank's share of the per-plane hot path.  ``dist_group`` = None for a single GPU, otherwise an initialised
    torch.distributed process group (NCCL) with one rank per GPU."""

    def __init__(self, sht_order, lmax=None, ray_order=None, ring_weights=None, dist_group=None, device=None, fused=True,
                 halo_deg=1.0, overlap_rays=False):
        import torch.distributed as dist
        self.dist = dist if dist_group is not None else None
        self.group = dist_group
        self.nranks = dist.get_world_size(dist_group) if dist_group is not None else 1
        self.rank = dist.get_rank(dist_group) if dist_group is not None else 0
        self.plan = sht.HEALPixSHTPlan(sht_order, lmax, ring_weights, self.nranks, self.rank, device=device)
        self.device = self.plan.device
        self.lib = self.plan.lib
        self.order = int(sht_order)
        self.npix = self.plan.npix
        self.ray_order = self.order if ray_order is None else int(ray_order)
        p = self.plan
        f64 = dict(dtype=torch.float64, device=self.device)
        self.g_send = torch.empty(2 * max(p.g_send_total, 1), **f64)
        self.b_send = torch.empty(2 * max(p.b_send_total, 1), **f64)
        if self.nranks > 1:
            self.g_recv = torch.empty(2 * max(p.g_recv_total, 1), **f64)
            self.b_recv = torch.empty(2 * max(p.b_recv_total, 1), **f64)
        else:
            self.g_recv, self.b_recv = self.g_send, self.b_send
        self.alm_re = torch.empty(max(p.Nlm, 1), **f64)
        self.alm_im = torch.empty(max(p.Nlm, 1), **f64)
        self.maps = None
        self.summary = torch.zeros(6, **f64)
        self._dens = [torch.zeros(self.npix, dtype=torch.float32, device=self.device) for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=self.device, priority=-1)   # runs behind the compute kernels, gets SM slots first
        self._staged = None        # (host map, scalings, buffer index, ready event) of a prefetched plane
        self._dens_free = [None, None]   # event after the last kernel that read each density buffer
        # overlap_rays: the ray kernel of plane p runs on its own stream while the next plane's density load, ring FFT and
        # Legendre analysis proceed (the SHT of plane p+1 does not depend on the rays); the next plane's ring synthesis,
        # which overwrites the maps, waits for it.  The ray kernel (128 registers) fits beside two Legendre CTAs on an SM.
        self.overlap_rays = bool(overlap_rays)
        self._ray_stream = torch.cuda.Stream(device=self.device)
        self._rays_done = None
        self.ray_events = []
        self.fused = False
        self.halo_deg = float(halo_deg)
        self.coarse_order = 5
        self._need = None          # device mask of the coarse cells each rank needs (fused exchange only)
        self._err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._peer_bufs = []
        if self.nranks > 1 and fused:
            self._setup_peer_exchange()
        if self.maps is None:
            self.maps = torch.zeros((6, self.npix), dtype=torch.float32, device=self.device)
        self.rays = None
        self.nrays = 0
        self.first_nest = 0

    # ---- fused exchange over peer memory ----
    def _dev_view(self, ptr, shape, dtype):
        """torch tensor over device memory owned by the library (CUDA array interface, zero copy)."""
        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": {torch.float64: "<f8", torch.float32: "<f4"}[dtype],
                                      "data": (int(ptr), False), "version": 2, "strides": None}
        return torch.as_tensor(h, device=self.device)

    def _setup_peer_exchange(self):
        """Allocate the receive buffers as peer-mappable memory, exchange the IPC handles through the process group
        and hand the peer pointers to the plan (clb_sht_plan_set_peers).  Falls back to the NCCL exchange, on all
        ranks together, if any mapping fails."""
        p, L = self.plan, self.lib
        sizes = [16 * max(p.g_send_total, 1), 16 * max(p.b_recv_total, 1), 4 * 6 * self.npix]   # g send, b receive, maps
        own = [L.clb_peer_alloc(n) for n in sizes]
        handles = []
        for ptr in own:
            buf = C.create_string_buffer(64)
            L.clb_peer_export(ptr, buf)
            handles.append(buf.raw)
        gathered = [None] * self.nranks
        self.dist.all_gather_object(gathered, handles, group=self.group)
        peers = [[None] * 3 for _ in range(self.nranks)]
        ok = 1
        for q in range(self.nranks):
            for k in range(3):
                if q == self.rank:
                    peers[q][k] = own[k]
                else:
                    ptr = L.clb_peer_import(C.create_string_buffer(gathered[q][k], 64))
                    if not ptr:
                        ok = 0
                    peers[q][k] = ptr
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        self._peer_bufs = (own, peers)
        if int(flag.item()) == 0:
            return   # stay on the NCCL path (buffers are simply not used)
        g_arr = (C.c_void_p * self.nranks)(*[peers[q][0] for q in range(self.nranks)])
        b_arr = (C.c_void_p * self.nranks)(*[peers[q][1] for q in range(self.nranks)])
        L.clb_sht_plan_set_peers(p._h, g_arr, b_arr)
        self._peer_maps = (C.c_void_p * (6 * self.nranks))(*[peers[q][2] + 4 * self.npix * k for q in range(self.nranks) for k in range(6)])
        self.g_send = self._dev_view(own[0], (2 * max(p.g_send_total, 1),), torch.float64)
        self.g_recv = None         # analysis reads g from the ring owners' send buffers
        self.b_recv = self._dev_view(own[1], (2 * max(p.b_recv_total, 1),), torch.float64)
        self.maps = self._dev_view(own[2], (6, self.npix), torch.float32)
        self.maps.zero_()
        self.b_send = None         # synthesis stores b into the ring owners' receive buffers
        self._tiny = torch.zeros(1, dtype=torch.float32, device=self.device)
        # halo-limited map broadcast: a pixel goes to the ranks whose ray domain, grown by halo_deg, can reach it.
        # Cells of a coarse NEST grid stand in for the reference's halo bundle cells (raytrace_utils.c:116-161); the
        # margin adds two coarse cell radii (a HEALPix pixel's radius is < 1.2 x its mean spacing).
        if self.halo_deg > 0 and self.order >= self.coarse_order:
            spacing = math.sqrt(4.0 * math.pi / (12 << (2 * self.coarse_order)))
            margin = math.radians(self.halo_deg) + 2.0 * 1.2 * spacing
            mask = np.zeros(12 << (2 * self.coarse_order), dtype=np.uint8)
            L.clb_domain_masks(self.ray_order, self.nranks, self.coarse_order, margin, mask.ctypes.data)
            self._need = torch.from_numpy(mask).to(self.device)
            self.need_fraction = float(np.unpackbits(mask[:, None], axis=1)[:, -self.nranks:].sum()) / (mask.size * self.nranks)
        self.fused = True
        self.dist.barrier(group=self.group)

    def close(self):
        """Release the peer mappings and the library-owned buffers (after every rank has stopped using them)."""
        self._wait_rays()
        if self._peer_bufs:
            torch.cuda.synchronize()
            if self.nranks > 1:
                self.dist.barrier(group=self.group)
            own, peers = self._peer_bufs
            self.maps = self.g_send = self.b_recv = None
            for q in range(self.nranks):
                if q != self.rank:
                    for ptr in peers[q]:
                        if ptr:
                            self.lib.clb_peer_release(ptr)
            if self.nranks > 1:
                self.dist.barrier(group=self.group)
            for ptr in own:
                self.lib.clb_peer_free(ptr)
            self._peer_bufs = []
            self.fused = False

    def _stream_barrier(self):
        """All ranks' work enqueued so far has completed before anything enqueued afterwards starts (stream ordered)."""
        self.dist.all_reduce(self._tiny, group=self.group)

    # ---- rays ----
    def init_rays(self, binL_2):
        """alloc_rays + init_rays (raytrace_utils.c:265-347) for this rank's contiguous NEST range."""
        self._wait_rays()
        nray_tot = 12 << (2 * self.ray_order)
        lo = (nray_tot * self.rank) // self.nranks
        hi = (nray_tot * (self.rank + 1)) // self.nranks
        self.first_nest, self.nrays = lo, hi - lo
        self.rays = torch.empty(max(self.nrays, 1) * 176, dtype=torch.uint8, device=self.device)
        self.lib.clb_ray_init_dev(self.rays.data_ptr(), self.nrays, lo, self.ray_order, float(binL_2), self._stream())
        return self.nrays

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def _all_to_all(self, send, recv, send_counts, recv_counts):
        if self.nranks == 1:
            return send
        self.dist.all_to_all_single(recv[:2 * sum(recv_counts)], send[:2 * sum(send_counts)],
                                    output_split_sizes=[2 * c for c in recv_counts],
                                    input_split_sizes=[2 * c for c in send_counts], group=self.group)
        return recv

    # ---- the SHT Poisson solve: counts map in self.maps[0] -> six derivative maps in self.maps ----
    def load_density(self, counts_map, premul, densmul, backdens, dst=None):
        """Scale this rank's rings of a full-sky count map (device tensor or pinned host tensor, RING float32) into the
        density buffer: shtpoissonsolve.c:342-502 for the raw-map input path."""
        assert counts_map.dtype == torch.float32 and counts_map.numel() == self.npix and counts_map.is_contiguous()
        assert counts_map.is_cuda or counts_map.is_pinned(), "host maps must be pinned (they are read by the GPU directly)"
        dst = self._dens[0] if dst is None else dst
        self.lib.clb_load_density_dev(self.plan._h, counts_map.data_ptr(), dst.data_ptr(), float(premul), float(densmul),
                                      float(backdens), self._stream())
        return dst

    def solve(self, density=None, mark=None):
        """density map (this rank's rings valid; default: the buffer load_density filled) -> six derivative maps.
        ``mark(name)`` (optional) is called after each stage has been enqueued (bench.py records CUDA events there)."""
        mark = mark or (lambda name: None)
        p = self.plan
        dens = self._dens[0] if density is None else density
        if self.fused:
            # producers store into the consumers' buffers over NVLink; barriers order producer and consumer stages
            self._stream_barrier()   # every rank is done with the previous plane's g, b and maps
            p.ring_analysis(dens, self.g_send); mark("fft_analysis")
            self._stream_barrier(); mark("a2a_g")
            self.lib.clb_legendre_analysis_dev(p._h, None, self.alm_re.data_ptr(), self.alm_im.data_ptr(), 1, self._stream())
            mark("legendre_analysis")
            self.lib.clb_legendre_synthesis_dev(p._h, self.alm_re.data_ptr(), self.alm_im.data_ptr(), None, self._stream())
            mark("legendre_synthesis")
            self._wait_rays()        # the previous plane's ray kernel still reads the maps (every rank, before the barrier)
            self._stream_barrier(); mark("a2a_b")
            p.ring_synthesis(self.b_recv, self.maps); mark("fft_synthesis")
            ptrs = (C.c_void_p * 6)(*[self.maps[k].data_ptr() for k in range(6)])
            self.lib.clb_maps_broadcast_dev(p._h, ptrs, self._peer_maps, None if self._need is None else self._need.data_ptr(),
                                            self.coarse_order, self._stream())
            self._stream_barrier(); mark("map_allreduce")
            return self.maps
        p.ring_analysis(dens, self.g_send); mark("fft_analysis")
        g = self._all_to_all(self.g_send, self.g_recv, p.counts[0], p.counts[1]); mark("a2a_g")
        p.legendre_analysis(g, self.alm_re, self.alm_im, poisson_filter=True); mark("legendre_analysis")
        p.legendre_synthesis(self.alm_re, self.alm_im, self.b_send); mark("legendre_synthesis")
        b = self._all_to_all(self.b_send, self.b_recv, p.counts[2], p.counts[3]); mark("a2a_b")
        self._wait_rays()
        if self.nranks > 1:
            self.maps.zero_()
        p.ring_synthesis(b, self.maps); mark("fft_synthesis")
        if self.nranks > 1:
            self.dist.all_reduce(self.maps, group=self.group)   # disjoint ring sets: x + 0 is exact
        mark("map_allreduce")
        return self.maps

    def _wait_rays(self):
        """The current stream waits for a ray kernel still running on the ray stream (overlap_rays)."""
        if self._rays_done is not None:
            torch.cuda.current_stream().wait_event(self._rays_done)
            self._rays_done = None

    def alm2allmaps(self, alm_re, alm_im):
        """alm2allmaps_mpi over the ranks of the group: local alm (this rank's m) -> the six full maps on every rank."""
        p = self.plan
        self._wait_rays()
        if self.fused:
            self._stream_barrier()
            self.lib.clb_legendre_synthesis_dev(p._h, alm_re.data_ptr(), alm_im.data_ptr(), None, self._stream())
            self._stream_barrier()
            p.ring_synthesis(self.b_recv, self.maps)
            ptrs = (C.c_void_p * 6)(*[self.maps[k].data_ptr() for k in range(6)])
            self.lib.clb_maps_broadcast_dev(p._h, ptrs, self._peer_maps, None if self._need is None else self._need.data_ptr(),
                                            self.coarse_order, self._stream())
            self._stream_barrier()
            return self.maps
        p.legendre_synthesis(alm_re, alm_im, self.b_send)
        b = self._all_to_all(self.b_send, self.b_recv, p.counts[2], p.counts[3])
        if self.nranks > 1:
            self.maps.zero_()
        p.ring_synthesis(b, self.maps)
        if self.nranks > 1:
            self.dist.all_reduce(self.maps, group=self.group)
        return self.maps

    def ray_update(self, wpp1, wp, wpm1, with_summary=False):
        """zero + interpolate + propagate: rayprop_sphere(planeRadPlus1, planeRad, planeRadMinus1) as called at
        raytrace.c:262, preceded by the reset of raytrace.c:213-230 and the interpolation of shtpoissonsolve.c:666-702.
        with_summary: also accumulate the six plane sums into self.summary (same kernel, no second pass)."""
        ptrs = (C.c_void_p * 6)(*[self.maps[k].data_ptr() for k in range(6)])
        need = None if self._need is None else self._need.data_ptr()

        def launch(stream):
            self.lib.clb_ray_step_ex_dev(self.rays.data_ptr(), self.nrays, ptrs, self.order, float(wpp1), float(wp), float(wpm1),
                                         MODE_ZERO | MODE_INTERP | MODE_PROP, need, self.coarse_order if need else 0, self.rank,
                                         self._err.data_ptr() if need else None,
                                         self.summary.data_ptr() if with_summary else None, stream)
        if not self.overlap_rays:
            launch(self._stream())
            return
        self._wait_rays()                      # planes are sequential for the rays
        maps_ready = torch.cuda.Event(); maps_ready.record()
        self._ray_stream.wait_event(maps_ready)
        t0 = torch.cuda.Event(enable_timing=True); t0.record(self._ray_stream)
        launch(self._ray_stream.cuda_stream)
        done = torch.cuda.Event(enable_timing=True); done.record(self._ray_stream)
        self._rays_done = done
        self.ray_events = (self.ray_events + [(t0, done)])[-64:]     # duration of the overlapped kernel, for bench.py

    def sync_rays(self):
        """Join the ray stream (before reading rays, summaries, or timing)."""
        self._wait_rays()

    def check_halo(self):
        """Raise if a ray left the part of the sky this rank receives (the reference aborts on a missing map cell,
        shtpoissonsolve.c:683-689); synchronises."""
        if self._need is not None and int(self._err.item()) != 0:
            raise RuntimeError("calclens_b200: a ray left its domain + halo (%.2f deg); raise halo_deg" % self.halo_deg)

    def prefetch(self, counts_map, premul, densmul, backdens):
        """Start loading the NEXT plane's density on the copy stream while the current plane computes: the host map is
        read by the GPU directly (only this rank's rings), scaled, and parked in the spare density buffer."""
        k = 1 if (self._staged is None or self._staged[2] == 0) else 0
        cs = self._copy_stream
        if self._dens_free[k] is not None:
            cs.wait_event(self._dens_free[k])
        with torch.cuda.stream(cs):
            self.load_density(counts_map, premul, densmul, backdens, dst=self._dens[k])
            ev = torch.cuda.Event(); ev.record(cs)
        self._next_staged = (counts_map, (float(premul), float(densmul), float(backdens)), k, ev)

    def step(self, counts_map, premul, densmul, backdens, wpp1, wp, wpm1, read_summary=True, prefetch=None):
        """One lens plane.  ``counts_map``: RING float32 full-sky map, a device tensor or a pinned host tensor.
        ``prefetch`` = (next_counts_map, premul, densmul, backdens) starts the next plane's load behind this plane's
        kernels; a later step() given that same map picks the staged density up instead of loading again."""
        st = self._staged
        key = (float(premul), float(densmul), float(backdens))
        if st is not None and st[0] is counts_map and st[1] == key:
            torch.cuda.current_stream().wait_event(st[3])
            k = st[2]
        else:
            k = 0 if st is None else 1 - st[2]
            self.load_density(counts_map, premul, densmul, backdens, dst=self._dens[k])
            self._staged = (None, None, k, None)
        self.solve(self._dens[k])
        ev = torch.cuda.Event(); ev.record()
        self._dens_free[k] = ev
        self._staged = (None, None, k, None)
        self.ray_update(wpp1, wp, wpm1, with_summary=read_summary)
        if prefetch is not None:
            self.prefetch(*prefetch)
            self._staged = self._next_staged
        if not read_summary:
            return None
        self._wait_rays()
        if self.nranks > 1:
            self.dist.all_reduce(self.summary, group=self.group)
        out = self.summary.cpu().numpy()
        self.check_halo()
        return out

    def rays_host(self):
        from .rays import rays_from_device
        self._wait_rays()
        return rays_from_device(self.rays[:self.nrays * 176])
