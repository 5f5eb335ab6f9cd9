"""CPU-side multi-rank test (gloo, world_size 2): the two SHT transposes as the N>1 path runs them --
torch.distributed.all_to_all_single with the split sizes of the exchange layout."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, order, lmax, out):
    sys.path.insert(0, ROOT)
    import calclens_b200 as clb
    from calclens_b200 import layout
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nside = 1 << order
    rp_owner, m_owner = clb.default_owners(order, lmax, world)
    L = layout.ExchangeLayout(nside, lmax, world, rank, rp_owner, m_owner)
    ok = True
    # analysis transpose: complex doubles as pairs of float64
    send = torch.zeros(2 * sum(L.g_send_counts), dtype=torch.float64)
    for m in range(lmax + 1):
        for rp in L.my_rp:
            for h in (0, 1):
                i = L.g_send_index(m, rp, h)
                send[2 * i] = float(m); send[2 * i + 1] = float(2 * rp + h)
    recv = torch.full((2 * sum(L.g_recv_counts),), -1.0, dtype=torch.float64)
    dist.all_to_all_single(recv, send, output_split_sizes=[2 * c for c in L.g_recv_counts],
                           input_split_sizes=[2 * c for c in L.g_send_counts])
    for m in L.my_m:
        for rp in range(2 * nside):
            for h in (0, 1):
                i = L.g_recv_index(m, rp, h)
                ok &= (recv[2 * i].item() == float(m)) and (recv[2 * i + 1].item() == float(2 * rp + h))
    # synthesis transpose
    send = torch.zeros(2 * sum(L.b_send_counts), dtype=torch.float64)
    for m in L.my_m:
        for f in range(6):
            for rp in range(2 * nside):
                for h in (0, 1):
                    i = L.b_send_index(m, f, rp, h)
                    send[2 * i] = float(m * 6 + f); send[2 * i + 1] = float(2 * rp + h)
    recv = torch.full((2 * sum(L.b_recv_counts),), -1.0, dtype=torch.float64)
    dist.all_to_all_single(recv, send, output_split_sizes=[2 * c for c in L.b_recv_counts],
                           input_split_sizes=[2 * c for c in L.b_send_counts])
    for m in range(lmax + 1):
        for f in range(6):
            for rp in L.my_rp:
                for h in (0, 1):
                    i = L.b_recv_index(m, f, rp, h)
                    ok &= (recv[2 * i].item() == float(m * 6 + f)) and (recv[2 * i + 1].item() == float(2 * rp + h))
    # map replication: disjoint ring sets summed over ranks reproduce the full map
    npix = 12 * nside * nside
    full = torch.arange(npix, dtype=torch.float32)
    mine = torch.zeros(npix, dtype=torch.float32)
    for rp in L.my_rp:
        r = rp + 1
        n = 4 * min(r, nside)
        start = 2 * r * (r - 1) if r < nside else 2 * nside * (nside - 1) + (r - nside) * 4 * nside
        mine[start:start + n] = full[start:start + n]
        if r != 2 * nside:
            s = npix - start - n
            mine[s:s + n] = full[s:s + n]
    dist.all_reduce(mine)
    ok &= bool(torch.equal(mine, full))
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(t.item() == 1.0))
    dist.destroy_process_group()


def test_transposes_over_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, 14, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
