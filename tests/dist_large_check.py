"""Property check of the sharded SHT at sizes no CPU oracle reaches, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
        tests/dist_large_check.py [order] [lmax]
A band-limited field with random a_lm (power-law spectrum) is synthesised over the ranks (fused exchange); on the rings
each rank owns it must satisfy the Laplacian identity grad_tt + grad_pp = synthesis of -l(l+1) a_lm, and analysing phi
must return the a_lm to HEALPix quadrature accuracy."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from calclens_b200 import poisson  # noqa: E402


def main():
    order = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2 << order
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    s = poisson.LensPlaneSolver(order, lmax, order, dist_group=dist.group.WORLD, device=local_rank, halo_deg=0.0)
    p = s.plan
    dev = s.device
    gen = torch.Generator(device=dev); gen.manual_seed(11 + rank)
    mloc = torch.as_tensor(np.asarray(p.m_local), device=dev, dtype=torch.int64)
    cnt = lmax + 1 - mloc
    start = torch.cumsum(cnt, 0) - cnt
    ls = (torch.arange(int(cnt.sum()), device=dev, dtype=torch.int64) - torch.repeat_interleave(start, cnt)
          + torch.repeat_interleave(mloc, cnt)).double()
    ms = torch.repeat_interleave(mloc, cnt)
    amp = (ls + 10.0) ** -1.1
    are = torch.randn(p.Nlm, generator=gen, device=dev, dtype=torch.float64) * amp
    aim = torch.randn(p.Nlm, generator=gen, device=dev, dtype=torch.float64) * amp
    aim[ms == 0] = 0.0
    are[ls == 0] = 0.0
    maps = s.alm2allmaps(are, aim).clone()
    lap = s.alm2allmaps(-ls * (ls + 1) * are, -ls * (ls + 1) * aim)[0].clone()
    dens = maps[0].contiguous()   # analysing phi only reads this rank's own rings
    # Laplacian identity on every pixel this rank holds (halo_deg=0 -> full broadcast, so all of them)
    num = (maps[3].double() + maps[5].double() - lap.double()).pow(2).sum()
    den = lap.double().pow(2).sum()
    rel = float((num / den).sqrt())
    # round trip: map2alm of phi (fused exchange of g)
    s._stream_barrier()
    p.ring_analysis(dens, s.g_send)
    s._stream_barrier()
    bre, bim = s.alm_re, s.alm_im
    s.lib.clb_legendre_analysis_dev(p._h, None, bre.data_ptr(), bim.data_ptr(), 0, s._stream())
    e = torch.stack([((bre - are).pow(2) + (bim - aim).pow(2)).sum(), (are.pow(2) + aim.pow(2)).sum()])
    dist.all_reduce(e)
    err = float((e[0] / e[1]).sqrt())
    ok = rel < 2e-5 and err < 5e-3
    if rank == 0:
        print("order %d lmax %d world %d: Laplacian identity rel L2 %.3e, round-trip alm rel L2 %.3e -> %s"
              % (order, lmax, world, rel, err, "OK" if ok else "FAIL"))
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
