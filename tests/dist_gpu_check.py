"""Multi-GPU check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
Every rank runs the sharded lens-plane step; rank 0 also runs the single-GPU step and the results must agree bit for
bit for the maps and alm-derived quantities, and to 1e-12 for the ray sums (different summation order only)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from calclens_b200 import poisson  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    order, lmax, ray_order = 7, 256, 7
    npix = 12 << (2 * order)
    rng = np.random.default_rng(77)
    counts = torch.from_numpy((8.0 * rng.lognormal(sigma=0.5, size=npix)).astype(np.float32)).pin_memory()
    premul, densmul, backdens = np.float32(1.0), np.float32(2e-4), np.float32(8.0 * np.exp(0.125) * 2e-4)
    fused = os.environ.get("CLB_FUSED", "1") != "0"
    halo = float(os.environ.get("CLB_HALO", "1.0"))
    solver = poisson.LensPlaneSolver(order, lmax, ray_order, dist_group=dist.group.WORLD, device=local_rank, fused=fused, halo_deg=halo)
    if rank == 0:
        print("exchange:", "fused peer stores" if solver.fused else "NCCL all-to-all")
    if fused and not solver.fused:
        print("WARNING: fused exchange requested but peer mapping failed")
    solver.init_rays(15.0)
    planes = [(45.0, 15.0, 0.0), (75.0, 45.0, 15.0), (105.0, 75.0, 45.0), (135.0, 105.0, 75.0), (165.0, 135.0, 105.0)]
    pm = [np.float32(premul * (1.0 + 0.25 * k)) for k in range(len(planes))]     # a different shell per plane
    # two planes per SHT pass (CLB_PAIR=1, fused solver only): planes (0,1) and (2,3) share their Legendre passes
    pairs = solver.fused and os.environ.get("CLB_PAIR", "1") != "0"
    sums = []
    for k, pl in enumerate(planes):
        pair = (counts, pm[k + 1], densmul, backdens) if (pairs and k in (0, 2)) else None
        sums.append(solver.step(counts, pm[k], densmul, backdens, *pl, pair=pair))
    if rank == 0:
        print("two planes per SHT pass:", bool(pairs))
    maps_d = solver.maps.cpu().numpy()
    rays_d = solver.rays_host().copy()
    gathered = [None] * world
    dist.all_gather_object(gathered, (solver.first_nest, rays_d.tobytes()))
    ok = True
    if rank == 0:
        single = poisson.LensPlaneSolver(order, lmax, ray_order, device=local_rank)
        single.init_rays(15.0)
        sums1 = [single.step(counts, pm[k], densmul, backdens, *pl) for k, pl in enumerate(planes)]
        maps_s = single.maps.cpu().numpy()
        rays_s = single.rays_host()
        if solver._need is None:
            ok &= bool(np.array_equal(maps_d, maps_s))
            print("maps identical:", np.array_equal(maps_d, maps_s))
        else:   # halo-limited broadcast: this rank holds only the part of the sky its rays can reach
            same_frac = float((maps_d == maps_s).mean())
            print("halo-limited broadcast: rank 0 holds %.1f%% of the pixels (needs %.1f%% of the coarse cells)"
                  % (100 * same_frac, 100 * float((solver._need.cpu().numpy() & 1).mean())))
        allrays = b"".join(x[1] for x in sorted(gathered))
        same = allrays == rays_s.tobytes()
        print("rays identical:", same, "nrays", rays_s.size)
        ok &= same
        for a, b in zip(sums, sums1):
            rel = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
            ok &= rel < 1e-12
        print("ray sums rel diff ok:", ok, sums[-1][:3], sums1[-1][:3])
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(t, 0)
    solver.close()
    dist.destroy_process_group()
    if t.item() != 1.0:
        sys.exit(1)
    if rank == 0:
        print("DIST CHECK OK world=%d" % world)


if __name__ == "__main__":
    main()
