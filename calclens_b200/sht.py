"""Host-side mirror of the reference's SHT interface (healpix_shtrans.h) over the CUDA library.

Names follow the reference: ``healpixsht_plan`` -> :class:`HEALPixSHTPlan`, ``map2alm_mpi``, ``alm2allmaps_mpi``,
``read_ring_weights``.  Arrays on the device are torch tensors (torch is used for memory, streams and
torch.distributed only); the arithmetic is in libcalclens_b200.so.
"""
import ctypes as C
import os
import struct

import numpy as np
import torch

from . import _lib


def order2nside(order):
    return 1 << order


def order2npix(order):
    return 12 << (2 * order)


def order2lmax(order):
    """The reference's default band limit, healpix_shtrans.c:518-521."""
    return 3 * order2nside(order) - 1


def num_lms(lmax):
    """healpix_shtrans.c:528-531."""
    return (lmax + 1) * (lmax + 1) - lmax * (lmax + 1) // 2


def lm2index(l, m, lmax):
    """healpix_shtrans.c:523-526 (single rank, m-major)."""
    return (m + 1) * (lmax + 1) - m * (m + 1) // 2 - (lmax - m + 1) + l - m


def read_ring_weights(path, order):
    """HEALPix ring weights, first BINTABLE extension, column 1, big-endian doubles
    (the file read_ring_weights opens, healpix_shtrans.c:361-423: '<path>/weight_ring_n%05d.fits')."""
    nside = order2nside(order)
    fname = os.path.join(path, "weight_ring_n%05d.fits" % nside)
    with open(fname, "rb") as f:
        data = f.read()

    def header(off):
        cards = {}
        while True:
            block = data[off:off + 2880]
            off += 2880
            end = False
            for i in range(36):
                card = block[80 * i:80 * i + 80].decode("ascii", "replace")
                key = card[:8].strip()
                if key == "END":
                    end = True
                    break
                if card[8:10] == "= ":
                    cards[key] = card[10:].split("/")[0].strip().strip("'").strip()
            if end:
                return cards, off
    _, off = header(0)
    cards, off = header(off)
    naxis1, naxis2 = int(cards["NAXIS1"]), int(cards["NAXIS2"])
    tform = cards["TFORM1"]
    repeat = int(tform[:-1]) if tform[:-1] else 1
    if tform[-1] != "D":
        raise ValueError("unexpected TFORM1 %r in %s" % (tform, fname))
    out = np.empty(naxis2 * repeat)
    for row in range(naxis2):
        out[row * repeat:(row + 1) * repeat] = struct.unpack(">%dd" % repeat, data[off + row * naxis1:off + row * naxis1 + 8 * repeat])
    if out.size != 2 * nside:
        raise ValueError("expected %d ring weights, found %d" % (2 * nside, out.size))
    return out


def default_owners(order, lmax, nranks):
    """Ring pairs are dealt to ranks round-robin in groups of four adjacent pairs (128-byte runs in the exchange
    buffers), m round-robin: every rank sees all latitudes and all m magnitudes, which balances both the FFT and the
    Legendre stage without the reference's cost polynomials (healpix_shtrans.c:219-250, :597-626)."""
    nrp = 2 * order2nside(order)
    group = 4 if nrp >= 32 * nranks else 1
    rp_owner = ((np.arange(nrp) // group) % nranks).astype(np.int32)
    m_owner = (np.arange(lmax + 1) % nranks).astype(np.int32)
    return rp_owner, m_owner


class HEALPixSHTPlan:
    """Counterpart of the reference's HEALPixSHTPlan (healpix_shtrans.h:30-45) for one rank / one GPU."""

    def __init__(self, order, lmax=None, ring_weights=None, nranks=1, rank=0, rp_owner=None, m_owner=None, device=None):
        self.lib = _lib.load()
        self.order = int(order)
        self.lmax = int(order2lmax(order) if lmax is None else lmax)
        self.nside = order2nside(order)
        self.npix = order2npix(order)
        self.nranks, self.rank = int(nranks), int(rank)
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib.clb_set_device(self.device.index)
        w = None if ring_weights is None else np.ascontiguousarray(ring_weights, dtype=np.float64)
        if w is not None and w.size != 2 * self.nside:
            raise ValueError("ring_weights must have 2*Nside entries")
        if nranks > 1 and (rp_owner is None or m_owner is None):
            rp_owner, m_owner = default_owners(order, self.lmax, nranks)
        ro = None if rp_owner is None else np.ascontiguousarray(rp_owner, dtype=np.int32)
        mo = None if m_owner is None else np.ascontiguousarray(m_owner, dtype=np.int32)
        self.ring_weights = w
        self._h = self.lib.clb_sht_plan_create(self.order, self.lmax, None if w is None else w.ctypes.data, self.nranks,
                                               self.rank, None if ro is None else ro.ctypes.data,
                                               None if mo is None else mo.ctypes.data)
        self._query()

    def _query(self):
        q = lambda k: self.lib.clb_sht_plan_query(self._h, k)
        self.Nlm = q(2)
        self.nrp_loc, self.nm_loc = q(3), q(4)
        self.g_send_total, self.g_recv_total, self.b_send_total, self.b_recv_total = q(5), q(6), q(7), q(8)
        self.counts = []
        for which in range(4):
            c = (C.c_long * self.nranks)()
            self.lib.clb_sht_plan_counts(self._h, which, c)
            self.counts.append([int(x) for x in c])
        ml = (C.c_int * max(self.nm_loc, 1))()
        self.lib.clb_sht_plan_local_m(self._h, ml)
        self.m_local = np.array(ml[:self.nm_loc], dtype=np.int64)
        rl = (C.c_int * max(self.nrp_loc, 1))()
        self.lib.clb_sht_plan_local_ring_pairs(self._h, rl)
        self.rp_local = np.array(rl[:self.nrp_loc], dtype=np.int64)

    def destroy(self):
        """healpixsht_destroy_plan, healpix_shtrans.c:496."""
        if getattr(self, "_h", None):
            self.lib.clb_sht_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ---- device stages (torch tensors on self.device) ----
    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def ring_analysis(self, map_dev, g_send=None):
        assert map_dev.dtype == torch.float32 and map_dev.is_cuda and map_dev.numel() == self.npix and map_dev.is_contiguous()
        if g_send is None:
            g_send = torch.empty(2 * max(self.g_send_total, 1), dtype=torch.float64, device=self.device)
        self.lib.clb_ring_analysis_dev(self._h, map_dev.data_ptr(), g_send.data_ptr(), self._stream())
        return g_send

    def legendre_analysis(self, g_recv, alm_re=None, alm_im=None, poisson_filter=False, nshell=1):
        """nshell = 2: two shells in one pass (g_recv holds them back to back, g_recv_total complex values apart; the alm
        come back the same way, Nlm apart)."""
        assert g_recv.dtype == torch.float64 and g_recv.is_cuda and g_recv.numel() >= 2 * nshell * self.g_recv_total
        if alm_re is None:
            alm_re = torch.empty(nshell * max(self.Nlm, 1), dtype=torch.float64, device=self.device)
            alm_im = torch.empty(nshell * max(self.Nlm, 1), dtype=torch.float64, device=self.device)
        assert alm_re.numel() >= nshell * self.Nlm and alm_im.numel() >= nshell * self.Nlm
        self.lib.clb_legendre_analysis_shells_dev(self._h, g_recv.data_ptr(), alm_re.data_ptr(), alm_im.data_ptr(),
                                                  1 if poisson_filter else 0, int(nshell), self._stream())
        return alm_re, alm_im

    def legendre_synthesis(self, alm_re, alm_im, b_send=None, nshell=1):
        assert alm_re.dtype == torch.float64 and alm_re.is_cuda and alm_im.is_cuda
        assert alm_re.numel() >= nshell * self.Nlm and alm_im.numel() >= nshell * self.Nlm
        if b_send is None:
            b_send = torch.empty(2 * nshell * max(self.b_send_total, 1), dtype=torch.float64, device=self.device)
        assert b_send.numel() >= 2 * nshell * self.b_send_total
        self.lib.clb_legendre_synthesis_shells_dev(self._h, alm_re.data_ptr(), alm_im.data_ptr(), b_send.data_ptr(), int(nshell),
                                                   self._stream())
        return b_send

    def ring_synthesis(self, b_recv, maps=None):
        assert b_recv.dtype == torch.float64 and b_recv.is_cuda and b_recv.numel() >= 2 * self.b_recv_total
        if maps is None:
            maps = torch.zeros((6, self.npix), dtype=torch.float32, device=self.device)
        assert maps.dtype == torch.float32 and maps.is_contiguous() and tuple(maps.shape) == (6, self.npix)
        ptrs = (C.c_void_p * 6)(*[maps[k].data_ptr() for k in range(6)])
        self.lib.clb_ring_synthesis_dev(self._h, b_recv.data_ptr(), ptrs, self._stream())
        return maps


# ---- the reference's two entry points, single rank, host arrays in and out ----
def map2alm_mpi(ringmap, plan, poisson_filter=False):
    """map2alm_mpi (map2alm_transpose_mpi.c:54) on a RING-ordered float32 host map -> (alm_real, alm_imag)."""
    m = np.ascontiguousarray(ringmap, dtype=np.float32)
    assert m.size == plan.npix
    are = np.empty(plan.Nlm); aim = np.empty(plan.Nlm)
    plan.lib.clb_map2alm(plan._h, m.ctypes.data, are.ctypes.data, aim.ctypes.data, 1 if poisson_filter else 0)
    return are, aim


def alm2allmaps_mpi(alm_real, alm_imag, plan):
    """alm2allmaps_mpi (alm2allmaps_transpose_mpi.c:53) -> float32 [6, Npix] RING maps
    (phi, grad_theta, grad_phi, grad_theta_theta, grad_theta_phi, grad_phi_phi)."""
    are = np.ascontiguousarray(alm_real, dtype=np.float64); aim = np.ascontiguousarray(alm_imag, dtype=np.float64)
    assert are.size == plan.Nlm and aim.size == plan.Nlm
    maps = np.empty((6, plan.npix), dtype=np.float32)
    plan.lib.clb_alm2allmaps(plan._h, are.ctypes.data, aim.ctypes.data, maps.ctypes.data)
    return maps
