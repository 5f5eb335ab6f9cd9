#!/usr/bin/env python
"""bench.py -- lens planes per second of the CALCLENS SHTONLY hot path on B200.

One "step" = one lens plane: density scaling -> map2alm -> Poisson filter -> alm2allmaps (six maps) -> per-ray
interpolation + propagation, i.e. do_healpix_sht_poisson_solve (shtpoissonsolve.c:38-708) plus the plane's
rayprop_sphere calls (raytrace.c:256-269), on synthetic lognormal count maps.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this framework (CUDA, sm_100a)
    python bench.py --impl reference [...]                        # the reference's own CPU code (oracle/_ref)

For N > 1 launch through torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

METRIC = "lens planes/sec (SHT Poisson+ray prop) at Nside 4096; ray-plane updates/s"


# ----------------------------------------------------------------------------------------------------------------
# algorithmic work
# ----------------------------------------------------------------------------------------------------------------
def triple_count(nside, lmax):
    """Number of (m, ring pair, l) triples with the reference's get_lmin_ylm cut (healpix_shtrans.c:533-544)."""
    npix = 12 * nside * nside
    r = np.arange(1, 2 * nside + 1, dtype=np.float64)
    fact2 = 4.0 / npix
    fact1 = 2 * nside * fact2
    cth = np.where(r < nside, 1.0 - r * r * fact2, (2 * nside - r) * fact1)
    sth = np.sqrt(np.maximum(0.0, (1.0 - cth) * (1.0 + cth)))
    total = 0
    m = np.arange(0, lmax + 1, dtype=np.float64)
    for lo in range(0, sth.size, 512):
        s = sth[lo:lo + 512][:, None]
        cut = np.trunc((m[None, :] - 40.0) / 1.35 / s)
        lmin = np.maximum(m[None, :], cut)
        total += int(np.maximum(0.0, lmax - lmin + 1).sum())
    return total


def work_model(nside, lmax, nrays):
    npix = 12 * nside * nside
    tri = triple_count(nside, lmax)
    rings = 4 * nside - 1
    return dict(
        tri=tri,
        # analysis: recurrence (DMUL + DFMA) + two accumulate DFMA per triple = 8 flop.  synthesis: SURVEY.md section 8d
        # counts 20 (recurrence + four complex sums); the three-sum formulation used here needs
        # 16, and that is the figure the roofline uses -- work that a better algorithm does not need is not "achieved"
        flops_analysis=8.0 * tri, flops_synthesis=16.0 * tri, flops_synthesis_survey=20.0 * tri,
        bytes_fft_analysis=4.0 * npix + 16.0 * rings * (lmax + 1),
        bytes_fft_synthesis=6.0 * (16.0 * rings * (lmax + 1) + 4.0 * npix),
        bytes_rays=448.0 * nrays)


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own MPI code on the host cores (oracle/cpu_arm.py)
# ----------------------------------------------------------------------------------------------------------------
def host_cores(sample_order):
    """Ranks for the CPU legs: the largest power of two <= host cores (the reference's hypercube exchange wants one),
    bounded by memory (one rank holds six maps plus its share of the rays)."""
    cores = os.cpu_count() or 1
    try:
        import psutil
        per_proc = 0.6e9 * 4 ** max(sample_order - 10, 0)
        cores = max(1, min(cores, int(0.5 * psutil.virtual_memory().available / per_proc)))
    except Exception:
        pass
    p = 1
    while 2 * p <= min(cores, 64):
        p *= 2
    return p


def cpu_reference_sample(sample_order, cores, target_nside, target_lmax, target_nrays, seed=0):
    """One lens plane of the reference on `cores` MPI-stub ranks at the sample size (same lmax/Nside ratio as the target),
    extrapolated to the target plane: Legendre time by the exact triple ratio, ring-FFT time by sum(n log2 n), ray time by
    the ray count.  Returns (planes/s at the target, detail dict, wall seconds)."""
    from oracle import cpu_arm
    s_nside = 1 << sample_order
    s_lmax = min(int(round(target_lmax * s_nside / float(target_nside))), 3 * s_nside - 1)
    return cpu_arm.sample(sample_order, s_lmax, sample_order, cores, seed, triple_count, (target_nside, target_lmax, target_nrays))


def cpu_sample_text(detail):
    m, e = detail["measured_seconds"], detail["extrapolation"]
    return ("reference map2alm_mpi + filter + alm2allmaps_mpi + shearinterp_comp + rayprop_sphere, %s built %s, %d ranks of the "
            "shared-memory MPI stub (real hypercube transposes; FFTW replaced by the FP64 shim), one plane at Nside=%d lmax=%d rays "
            "Nside=%d: %.2f s analysis, %.2f s synthesis, %.2f s rays (max over ranks); EXTRAPOLATED to the target plane: Legendre x%.1f "
            "(triples), ring FFT x%.1f (n log n), rays x%.1f -> %.1f s per plane" % (
                detail["library"], detail["flags"], detail["ranks"], detail["sample"]["nside"], detail["sample"]["lmax"],
                detail["sample"]["ray_nside"], m["map2alm_mpi"], m["filter+alm2allmaps_mpi"], m["shearinterp_comp+rayprop_sphere"],
                e["legendre_by_triples"], e["fft_by_npix_log_n"], e["rays_by_count"], detail["extrapolated_seconds_per_plane"]))


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    cfg = workload_config(a)
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcalclens_ref.so not built (needs /root/reference at build time)"}))
        return 0
    cores = host_cores(a.cpu_sample_order)
    vals, wall, detail = [], 0.0, None
    for i in range(a.warmup + a.steps):
        v, detail, w = cpu_reference_sample(a.cpu_sample_order, cores, a.nside, a.lmax, cfg["nrays"], seed=100 + i)
        if i >= a.warmup:
            vals.append(v); wall += w
    value = float(np.mean(vals)) if vals else 0.0
    line = {"metric": METRIC, "value": value, "unit": "planes/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 / value if value else None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": cfg["config"],
            "ray_plane_updates_per_s": value * cfg["nrays"],
            "cpu_baseline": {"value": value, "unit": "planes/s", "cores": cores, "kind": "reference", "sample": cpu_sample_text(detail),
                             "detail": detail},
            "e2e": {"value": value, "unit": "planes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
def workload_config(a):
    nrays = 12 * a.ray_nside * a.ray_nside
    return {"nrays": nrays,
            "config": {"workload": "SHTONLY synthetic lognormal shells Nside=%d lmax=%d, full-sky rays Nside=%d (%d rays), one lens plane per step"
                                   % (a.nside, a.lmax, a.ray_nside, nrays),
                       "nside": a.nside, "lmax": a.lmax, "ray_nside": a.ray_nside, "num_lens_planes": a.planes,
                       "l2_policy": "inputs larger than L2: every step streams a %.2f GB count map, %.1f GB of ray state and %.1f GB of "
                                    "ring/alm exchange buffers (L2 = 126 MB); distinct input maps are cycled"
                                    % (4e-9 * 12 * a.nside ** 2, 176e-9 * nrays, 16e-9 * 7 * (4 * a.nside) * (a.lmax + 1))}}


def _hash_normal(l, m, seed, which):
    """standard normals that depend only on (l, m, seed): the same alm on any number of ranks (32-bit integer hash of
    the indices, Box-Muller)"""
    import torch

    def h(salt):
        x = (l * 0x9E3779B1 + m * 0x85EBCA77 + (seed * 2 + salt) * 0xC2B2AE3D + which * 0x27D4EB2F) & 0xFFFFFFFF
        x = x ^ (x >> 16); x = (x * 0x7FEB352D) & 0xFFFFFFFF
        x = x ^ (x >> 15); x = (x * 0x846CA68B) & 0xFFFFFFFF
        x = x ^ (x >> 16)
        return (x.double() + 0.5) / 4294967296.0
    return torch.sqrt(-2.0 * torch.log(h(0))) * torch.cos(2.0 * math.pi * h(1))


def make_count_maps(solver, nmaps, lmax, seed=1234, nbar=8.0, sigma=0.5):
    """Synthetic lognormal over-density shells: Gaussian field with C_l ~ (l+10)^-1.2 synthesised
    with this library's own alm->map, delta = exp(sigma g - sigma^2/2) - 1, counts = nbar (1 + delta).  The alm depend only
    on (l, m, seed) and the normalisation is analytic, so every rank count gives the same shells.  Outside any timed region."""
    import torch
    p = solver.plan
    dev = solver.device
    ll = torch.arange(1, lmax + 1, dtype=torch.float64)
    var = float(((2 * ll + 1) / (4 * math.pi) * (ll + 10.0) ** -1.2).sum())     # field variance of the spectrum
    maps = []
    for k in range(nmaps):
        if p.nm_loc:
            mloc = torch.as_tensor(np.asarray(p.m_local), device=dev, dtype=torch.int64)
            cnt = lmax + 1 - mloc
            start = torch.cumsum(cnt, 0) - cnt
            ms = torch.repeat_interleave(mloc, cnt)
            li = torch.arange(int(cnt.sum()), device=dev, dtype=torch.int64) - torch.repeat_interleave(start, cnt) + ms
            ls = li.double()
            # a_lm = sqrt(C_l/2) (x + i y) for m > 0, sqrt(C_l) x for m = 0
            amp = torch.sqrt(0.5 * (ls + 10.0) ** -1.2)
            amp[ms == 0] *= math.sqrt(2.0)
            amp[li < 1] = 0.0
            are = _hash_normal(li, ms, seed + k, 0) * amp
            aim = _hash_normal(li, ms, seed + k, 1) * amp
            aim[ms == 0] = 0.0
        else:
            are = torch.zeros(1, device=dev, dtype=torch.float64); aim = torch.zeros(1, device=dev, dtype=torch.float64)
        solver.alm2allmaps(are.contiguous(), aim.contiguous())
        g = solver.maps[0].double() / math.sqrt(var)
        counts = (nbar * torch.exp(sigma * g - 0.5 * sigma * sigma)).float()
        maps.append(counts)
    return maps


def plane_schedule(first, n, shells, prefetch=False):
    """How a run of n consecutive planes starting at `first` is stepped: [(plane, partner or None, planes to prefetch)].
    Two shells per SHT pass (--shells 2, the default): plane 0 of the run is solved together with plane 1, plane 2 with plane 3,
    ...; an odd last plane is solved alone -- a run does exactly n planes of work, whatever was solved before or comes after.
    prefetch (host maps): the planes to come stream in behind the kernels of the step that frees the density buffers."""
    out = []
    for i in range(n):
        s = first + i
        partner = s + 1 if (shells == 2 and i % 2 == 0 and i + 1 < n) else None
        ahead = []
        if prefetch:
            ahead = [s + 1] if shells == 1 else ([s + 2, s + 3] if i % 2 == 0 else [])
        out.append((s, partner, ahead))
    return out


STAGES = ["scale", "fft_analysis", "a2a_g", "legendre_analysis", "legendre_synthesis", "a2a_b", "fft_synthesis", "map_allreduce", "rays"]


def measure(a, L, poisson, torch, dist, world, rank, local_rank, group, order, ray_order, lmax, steps, warmup, with_e2e=True):
    """device-resident and end-to-end planes/s of one configuration through clb_solver_* (calclens_b200/csrc/solver.cu)"""
    t_setup = time.time()
    solver = poisson.LensPlaneSolver(order, lmax, ray_order, dist_group=group, device=local_rank, fused=(a.exchange == "fused"))
    if world > 1 and a.exchange == "fused":
        assert solver.fused, "bench.py: the fused exchange is not active (peer mapping failed) -- refusing to report NCCL numbers as fused"
    cosmo = poisson.Cosmology(0.27)
    max_dist = 30.0 * a.planes
    pp = [poisson.plane_params(p, a.planes, max_dist, 0.27, cosmo) for p in range(a.planes)]
    solver.init_rays(pp[0]["binL"] / 2.0)
    nmaps = 3
    dev_maps = make_count_maps(solver, nmaps, lmax)
    host_maps = [torch.empty(solver.npix, dtype=torch.float32).pin_memory() for _ in range(nmaps)]
    for h, d in zip(host_maps, dev_maps):
        h.copy_(d)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup

    def plane_args(step, peek=False):
        if step % a.planes == 0 and step > 0 and not peek:
            solver.init_rays(pp[0]["binL"] / 2.0)   # a new light cone: rays back at the first shell
        p = pp[step % a.planes]
        binL = p["binL"]
        vshell = 4.0 * math.pi / 3.0 * ((p["wp"] + binL / 2) ** 3 - (p["wp"] - binL / 2) ** 3)
        part_mass = 0.27 * poisson.RHO_CRIT * vshell / (8.0 * solver.npix)   # the shell holds its mean mass
        premul, densmul, backdens = poisson.density_scalings(order, part_mass, p["densfact"], p["backdens"])
        return premul, densmul, backdens, p["wpp1"], p["wp"], p["wpm1"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=solver.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def run_planes(first, n, maps, read_summary=False, prefetch=False):
        for s, partner, ahead in plane_schedule(first, n, a.shells, prefetch):
            pair = None if partner is None else (maps[partner % nmaps],) + tuple(plane_args(partner, peek=True)[:3])
            pre = [(maps[q % nmaps],) + tuple(plane_args(q, peek=True)[:3]) for q in ahead] or None
            summ = solver.step(maps[s % nmaps], *plane_args(s), read_summary=read_summary, prefetch=pre, pair=pair)
            yield s, summ

    # ---- device-resident throughput ("value"): count maps already in HBM, no host synchronisation between planes
    for _ in run_planes(0, warmup, dev_maps):
        pass
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.clb_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()    # lets `ncu --profile-from-start off` list exactly the timed region's launches
    e0.record()
    for _ in run_planes(warmup, steps, dev_maps):
        pass
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    launches = L.clb_launch_count() - launches0
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / steps
    clocks = sampler.stop() if rank == 0 else None
    # per-stage times (CUDA events recorded inside the solver on the launching stream), a separate pass of the same steps;
    # with two shells per pass a pair's SHT stages are recorded on its first plane and the second records rays only, so the
    # sums divided by the number of planes are per-plane averages
    stage_ms = {k: 0.0 for k in STAGES}
    solver.set_timing(True)
    base = warmup + steps
    pass_ms = {1: [], 2: []}     # legendre synthesis launches by shells per pass (the roofline kernel), ms per launch
    for i, _ in enumerate(run_planes(base, steps, dev_maps)):
        st = solver.stage_ms()
        for k, v in zip(STAGES, st):
            stage_ms[k] += v / steps
        if st[4] > 0.0:     # (the second plane of a pair records no SHT stage)
            paired = a.shells == 2 and i % 2 == 0 and i + 1 < steps
            pass_ms[2 if paired else 1].append(st[4])
    solver.set_timing(False)
    base += steps
    out = {"ms_per_step": ms_per_step, "stage_ms": stage_ms, "pass_ms": pass_ms, "launches": launches, "clocks": clocks, "setup_s": t_setup,
           "npix": solver.npix, "fused": solver.fused, "host_barriers": getattr(solver, "host_barriers", False)}
    if with_e2e:
        # ---- end to end through the public API with host buffers: H2D of every plane's map (prefetched behind the previous
        # planes' kernels) + D2H of the six ray sums every step
        for _ in run_planes(base, 2, host_maps, read_summary=True, prefetch=True):
            pass
        base += 2
        barrier()
        e0.record()
        summ = None
        for _, summ in run_planes(base, steps, host_maps, read_summary=True, prefetch=True):
            pass
        e1.record()
        barrier()
        out["e2e_ms"] = max_over_ranks(e0.elapsed_time(e1)) / steps
        out["last_summary"] = [float(x) for x in summ]
    solver.close()
    return out


def run_sweep(a, L, poisson, torch, dist, world, rank, local_rank, group):
    """BASELINE configs[3]: SHT Poisson round trip (map2alm + filter + alm2allmaps) Nside 512..8192, lmax = 2 Nside, one JSON
    line per size with the CPU arm (measured directly up to the sample order, extrapolated above it) beside it."""
    import calclens_b200 as clb
    lines = []
    sizes = [512, 1024, 2048, 4096] + ([8192] if world >= 4 else [])
    for nside in sizes:
        order = int(round(math.log2(nside))); lmax = 2 * nside
        solver = poisson.LensPlaneSolver(order, lmax, order, dist_group=group, device=local_rank, fused=True)
        dens = torch.randn(solver.npix, device=solver.device, dtype=torch.float32)
        for _ in range(2):
            solver.solve(dens)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            solver.solve(dens)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=solver.device); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        solver.close()
        if rank == 0:
            tri = triple_count(nside, lmax)
            cpu = None
            if not a.no_cpu_baseline:
                try:
                    so = min(order, a.cpu_sample_order)
                    cores = host_cores(so)
                    v, detail, _ = cpu_reference_sample(so, cores, nside, lmax, 0)
                    cpu = {"sht_round_trips_per_s": v, "cores": cores, "kind": "reference", "extrapolated": so != order, "detail": detail}
                except Exception as exc:
                    cpu = {"failed": repr(exc)}
            lines.append({"metric": "SHT Poisson round trip (map2alm + filter + alm2allmaps) per second", "value": 1000.0 / ms, "unit": "round trips/s",
                          "ms_per_round_trip": ms, "n_gpus": world, "nside": nside, "lmax": lmax, "triples": tri,
                          "legendre_tflops_algorithmic": 24.0 * tri / (ms * 1e-3) / 1e12, "cpu_baseline": cpu, "steps": a.steps})
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nside", type=int, default=4096)
    ap.add_argument("--lmax", type=int, default=None)
    ap.add_argument("--ray-nside", type=int, default=None)
    ap.add_argument("--planes", type=int, default=50)
    ap.add_argument("--cpu-sample-order", type=int, default=10,
                    help="log2 Nside of the bounded CPU sample (the same in the reference arm and in the GPU arm's cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lmax3", action="store_true", help="skip the second measurement at the reference's own lmax = 3 Nside - 1")
    ap.add_argument("--shells", type=int, default=2, choices=[1, 2],
                    help="lens planes per SHT pass: 2 = consecutive planes share one pass of each Legendre kernel (the lambda_lm "
                         "recurrence is generated once for both), 1 = plane by plane like the reference")
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[3]: SHT round-trip sweep over Nside, one JSON line per size")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="multi-GPU exchange: stores into peer memory from the producing kernels, or NCCL all-to-all + all-reduce")
    a = ap.parse_args()
    if a.lmax is None:
        a.lmax = 2 * a.nside
    if a.ray_nside is None:
        a.ray_nside = a.nside
    if a.impl == "reference":
        return run_reference_arm(a)
    if a.exchange == "nccl" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        a.shells = 1     # the NCCL comparison arm runs plane by plane

    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner, ...) go to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from calclens_b200 import _lib, poisson

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            print("bench.py: --gpus %d needs torchrun with %d ranks" % (a.gpus, a.gpus), file=sys.stderr)
            return 2
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    L = _lib.load()
    order = int(round(math.log2(a.nside)))
    ray_order = int(round(math.log2(a.ray_nside)))
    cfg = workload_config(a)
    nrays_total = cfg["nrays"]

    if a.sweep:
        lines = run_sweep(a, L, poisson, torch, dist, world, rank, local_rank, group)
        if rank == 0:
            for ln in lines:
                real_stdout.write(json.dumps(ln) + "\n")
            real_stdout.flush()
        if world > 1:
            dist.destroy_process_group()
        return 0

    r = measure(a, L, poisson, torch, dist, world, rank, local_rank, group, order, ray_order, a.lmax, a.steps, a.warmup)
    ms_per_step, stage_ms = r["ms_per_step"], r["stage_ms"]
    value = 1000.0 / ms_per_step
    # second configuration: the band limit the reference itself runs at this Nside (healpix_shtrans.c:518-521)
    lmax3 = None
    if not a.no_lmax3 and a.lmax != 3 * a.nside - 1:
        r3 = measure(a, L, poisson, torch, dist, world, rank, local_rank, group, order, ray_order, 3 * a.nside - 1, max(2, a.steps // 2), 2,
                     with_e2e=False)
        lmax3 = {"lmax": 3 * a.nside - 1, "ms_per_step": r3["ms_per_step"], "value": 1000.0 / r3["ms_per_step"], "unit": "planes/s",
                 "stage_ms": r3["stage_ms"], "triples": triple_count(a.nside, 3 * a.nside - 1),
                 "note": "same workload at the reference's own hard-wired lmax = 3 Nside - 1 (SURVEY.md D1)"}

    if rank == 0:
        wm = work_model(a.nside, a.lmax, nrays_total)
        # FP64 peak: MEASURED_PEAKS.json has no FP64 entry, so measure the DFMA rate here (tools/fp64_peak) if the binary exists
        fp64_peak, fp64_src = 34.1, "DFMA micro-benchmark tools/fp64_peak on this pool's B200 (round 1 measurement, 34.1 TFLOP/s)"
        exe = os.path.join(HERE, "tools", "fp64_peak")
        if os.path.exists(exe) and world == 1:
            try:
                out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
                for ln in out.splitlines():
                    if ln.startswith("dfma sustained"):
                        fp64_peak = float(ln.split()[-2]); fp64_src = "DFMA micro-benchmark tools/fp64_peak run inside this bench (sustained)"
            except Exception:
                pass
        hbm_peak, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        try:
            mp_ = json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(mp_["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        # the dominant kernel: legendre_synthesis_kernel.  With two shells per pass one launch does two planes' worth of
        # algorithmic work; its duration is the CUDA-event time of that stage in the steps that ran a two-shell pass
        nsh = 2 if r["pass_ms"][2] else 1
        launch_ms = float(np.mean(r["pass_ms"][nsh])) if r["pass_ms"][nsh] else stage_ms["legendre_synthesis"]
        flops_launch = nsh * wm["flops_synthesis"] / world
        ach_syn = flops_launch / (launch_ms * 1e-3) / 1e12
        traffic = None
        try:
            tj = json.load(open(os.path.join(HERE, "profiles", "r02_traffic.json")))
            key = "legendre_synthesis_%dshell" % nsh
            if world == 1 and a.nside == 4096 and a.lmax == 8192 and key in tj:
                traffic = tj[key].get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": "legendre_synthesis_kernel<R,%d> (dominant; %d lens plane%s per launch)" % (nsh, nsh, "s" if nsh > 1 else ""),
                    "bound": "fp64", "achieved": ach_syn, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach_syn / fp64_peak, "traffic": traffic, "launch_ms": launch_ms,
                    "peak_source": fp64_src, "algorithmic_flops_per_launch": flops_launch,
                    "note": "FP64 FMA-pipe roofline (not tensor: B200's DMMA rate equals its DFMA rate and the recurrence is serial per ring). "
                            "algorithmic = 16 flop per (m, ring pair, l) triple of the reference's lmin cut and shell: 2-instruction recurrence + three "
                            "complex sums (DESIGN.md section 4) -- a two-shell pass EXECUTES 14 instead of 16 FP64 instructions per two triple-shells, "
                            "so a fraction close to 1 is the amortised recurrence, not a faster pipe; with SURVEY.md 8d's four-sum count of 20 the "
                            "same time gives %.2f TFLOP/s. traffic = dram bytes of one launch (ncu, profiles/r02_traffic.json)" % (
                                nsh * wm["flops_synthesis_survey"] / world / (launch_ms * 1e-3) / 1e12)}
        stages = {
            "legendre_analysis": {"bound": "fp64", "achieved": wm["flops_analysis"] / world / (stage_ms["legendre_analysis"] * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s"},
            "fft_analysis": {"bound": "hbm", "achieved": wm["bytes_fft_analysis"] / world / (stage_ms["fft_analysis"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"},
            "fft_synthesis": {"bound": "hbm", "achieved": wm["bytes_fft_synthesis"] / world / (stage_ms["fft_synthesis"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"},
            "rays": {"bound": "hbm", "achieved": wm["bytes_rays"] / world / (stage_ms["rays"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"},
        }
        for v in stages.values():
            v["frac"] = v["achieved"] / v["peak"]
        cpu_baseline = None
        if world == 1 and not a.no_cpu_baseline:
            try:
                from oracle import ref
                if ref.available():
                    cores = host_cores(a.cpu_sample_order)
                    v, detail, _ = cpu_reference_sample(a.cpu_sample_order, cores, a.nside, a.lmax, nrays_total)
                    cpu_baseline = {"value": v, "unit": "planes/s", "cores": cores, "kind": "reference", "sample": cpu_sample_text(detail),
                                    "detail": detail}
                else:
                    cpu_baseline = {"value": None, "unit": "planes/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
            except Exception as exc:   # the GPU numbers stand even if the CPU leg fails
                cpu_baseline = {"value": None, "unit": "planes/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (exc,)}
        line = {"metric": METRIC, "value": value, "unit": "planes/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg["config"],
                "ray_plane_updates_per_s": value * nrays_total,
                "e2e": {"value": 1000.0 / r["e2e_ms"], "unit": "planes/s", "ms_per_step": r["e2e_ms"], "h2d_bytes_per_step": 4 * r["npix"],
                        "d2h_bytes_per_step": 48 * world,
                        "api": "C ABI clb_solver_set_pair + clb_solver_set_next + clb_solver_step(pinned host count map, &sum6) of libcalclens_b200.so "
                               "(called through calclens_b200.poisson.LensPlaneSolver.step): host map in, six ray sums out, every plane; every rank "
                               "reads only its own rings of the host map",
                        "last_summary": r["last_summary"],
                        "checksum_note": "the six ray sums after the last timed plane; inputs depend only on (l, m, seed), so lines at different N are comparable"},
                "gpu_launches": int(r["launches"] * world),
                "shells_per_pass": a.shells,
                "clocks": r["clocks"],
                "roofline": roofline, "roofline_stages": stages, "stage_ms": stage_ms,
                "hbm_peak_source": hbm_src,
                "lmax_3nside_minus_1": lmax3,
                "exchange": (("fused peer loads/stores (CUDA IPC over NVLink) + device-side peer barrier" + (" [host barriers]" if r["host_barriers"] else ""))
                             if r["fused"] and world > 1 else "NCCL all-to-all-v + all-reduce" if world > 1 else "none (single GPU)"),
                "cpu_baseline": cpu_baseline,
                "setup_s": r["setup_s"]}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
