"""Exchange layouts of the two SHT transposes, host-side and CUDA-free.

The map side is sharded by ring pairs, the alm side by m.  Between the ring FFTs and the
Legendre stage the ranks exchange, per direction, one block per peer -- the counterpart of the sendcnts/recvcnts of
map2alm_transpose_mpi.c:329-347 (analysis: my rings x the peer's m) and alm2allmaps_transpose_mpi.c:656-672
(synthesis: my m x 6 fields x the peer's rings).  The C library computes the same tables for its kernels
(csrc/sht_plan.cu); tests check the two against each other, and the gloo tests run the exchange on CPU ranks.
All counts and offsets are in units of one complex double.
"""
import numpy as np


class ExchangeLayout:
    def __init__(self, nside, lmax, nranks, rank, rp_owner, m_owner):
        self.nside, self.lmax, self.nranks, self.rank = nside, lmax, nranks, rank
        nrp = 2 * nside
        self.rp_owner = np.asarray(rp_owner, dtype=np.int64)
        self.m_owner = np.asarray(m_owner, dtype=np.int64)
        assert self.rp_owner.size == nrp and self.m_owner.size == lmax + 1
        self.nrp_of = np.bincount(self.rp_owner, minlength=nranks)
        self.nm_of = np.bincount(self.m_owner, minlength=nranks)
        # index of every ring pair / m inside its owner's ascending list
        self.rp_local = np.zeros(nrp, dtype=np.int64)
        self.m_local = np.zeros(lmax + 1, dtype=np.int64)
        for q in range(nranks):
            sel = np.nonzero(self.rp_owner == q)[0]; self.rp_local[sel] = np.arange(sel.size)
            sel = np.nonzero(self.m_owner == q)[0]; self.m_local[sel] = np.arange(sel.size)
        self.my_rp = np.nonzero(self.rp_owner == rank)[0]
        self.my_m = np.nonzero(self.m_owner == rank)[0]
        nslot_mine = 2 * self.my_rp.size
        nm_mine = self.my_m.size
        self.nslot_mine = nslot_mine
        self.g_send_counts = [int(self.nm_of[q]) * nslot_mine for q in range(nranks)]
        self.g_recv_counts = [nm_mine * 2 * int(self.nrp_of[q]) for q in range(nranks)]
        self.b_send_counts = [nm_mine * 6 * 2 * int(self.nrp_of[q]) for q in range(nranks)]
        self.b_recv_counts = [int(self.nm_of[q]) * 6 * nslot_mine for q in range(nranks)]
        self.g_sbase = np.concatenate([[0], np.cumsum(self.g_send_counts)])
        self.g_rbase = np.concatenate([[0], np.cumsum(self.g_recv_counts)])
        self.b_sbase = np.concatenate([[0], np.cumsum(self.b_send_counts)])
        self.b_rbase = np.concatenate([[0], np.cumsum(self.b_recv_counts)])

    # analysis transpose --------------------------------------------------------------------------------------
    def g_send_index(self, m, rp, hemi):
        """where the FFT stage of the owner of ring pair rp (this rank) puts g_m of (rp, hemi)"""
        q = self.m_owner[m]
        return int(self.g_sbase[q] + self.m_local[m] * self.nslot_mine + 2 * self.rp_local[rp] + hemi)

    def g_recv_index(self, m, rp, hemi):
        """where the Legendre stage of the owner of m (this rank) finds g_m of (rp, hemi)"""
        q = self.rp_owner[rp]
        return int(self.g_rbase[q] + self.m_local[m] * 2 * self.nrp_of[q] + 2 * self.rp_local[rp] + hemi)

    # synthesis transpose -------------------------------------------------------------------------------------
    # Inside a peer block the order is [ring pair][field][m][hemisphere]: the ring FFT of a (ring, field) then reads its
    # b_m as one contiguous run in m (streaming HBM reads instead of one 16-byte gather per m at a stride of megabytes),
    # and the scatter sits on the Legendre side, whose 32-byte (north, south) stores the L2 merges into full lines.
    def b_send_index(self, m, field, rp, hemi):
        """where the Legendre stage of the owner of m (this rank) puts b of (field, rp, hemi)"""
        q = self.rp_owner[rp]
        return int(self.b_sbase[q] + ((self.rp_local[rp] * 6 + field) * self.my_m.size + self.m_local[m]) * 2 + hemi)

    def b_recv_index(self, m, field, rp, hemi):
        """where the FFT stage of the owner of ring pair rp (this rank) finds b_m of (field, rp, hemi)"""
        q = self.m_owner[m]
        return int(self.b_rbase[q] + ((self.rp_local[rp] * 6 + field) * int(self.nm_of[q]) + self.m_local[m]) * 2 + hemi)


    # fused exchange over peer memory (csrc/sht_plan.cu:sht_plan_set_peers) -------------------------------------
    def g_pull_index(self, m_idx, rp, hemi):
        """where THIS rank (owner of its m_idx-th m) reads g of (rp, hemi) inside the SEND buffer of rp's owner q:
        q's block for this rank starts after its blocks for the lower ranks (nm_of[r] * nslot_q elements each)."""
        q = self.rp_owner[rp]
        nslot_q = 2 * int(self.nrp_of[q])
        sbase = sum(int(self.nm_of[r]) * nslot_q for r in range(self.rank))
        return int(q), int(sbase + m_idx * nslot_q + 2 * self.rp_local[rp] + hemi)

    def b_push_index(self, m_idx, field, rp, hemi):
        """where THIS rank (owner of its m_idx-th m) stores b of (field, rp, hemi) inside the RECEIVE buffer of rp's
        owner q: q's block from this rank starts after the blocks from the lower ranks (nm_of[r] * 6 * nslot_q each)."""
        q = self.rp_owner[rp]
        nslot_q = 2 * int(self.nrp_of[q])
        rbase = sum(int(self.nm_of[r]) * 6 * nslot_q for r in range(self.rank))
        return int(q), int(rbase + ((self.rp_local[rp] * 6 + field) * self.my_m.size + m_idx) * 2 + hemi)


def ray_ranges(ray_order, nranks):
    """Contiguous NEST ranges of rays per rank = compact sky domains (cf. loadbalance.c:151-181 equal-area split)."""
    n = 12 << (2 * ray_order)
    return [((n * r) // nranks, (n * (r + 1)) // nranks) for r in range(nranks)]
