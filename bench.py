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
# reference arm / CPU baseline (the compiled reference functions in oracle/_ref on host cores)
# ----------------------------------------------------------------------------------------------------------------
def _cpu_sample_worker(args):
    order, lmax, ray_order, seed = args
    from oracle import ref
    rng = np.random.default_rng(seed)
    nside = 1 << order
    npix = 12 * nside * nside
    m = rng.lognormal(sigma=0.5, size=npix).astype(np.float32)
    m = (m * np.float32(8.0) - np.float32(8.0 * math.exp(0.125))).astype(np.float32)
    t0 = time.time()
    are, aim = ref.map2alm(order, lmax, m)
    t1 = time.time()
    are, aim = ref.poisson_filter(lmax, are, aim)
    maps = ref.alm2allmaps(order, lmax, are, aim)
    t2 = time.time()
    maps *= np.float32(1e-3 / max(float(np.abs(maps[3]).max()), 1e-30))
    rays = ref.init_rays(ray_order, 15.0)
    t3 = time.time()
    ref.shearinterp(order, min(order, 3), maps, rays)
    ref.rayprop(rays, 45.0, 15.0, 0.0)
    t4 = time.time()
    return (t1 - t0, t2 - t1, t4 - t3)


def cpu_reference_sample(sample_order, cores, target_nside, target_lmax, target_nrays, seed=0):
    """Run one bounded sample of the reference on `cores` processes in parallel (independent planes, which is how a
    user would occupy the cores without MPI) and extrapolate to the target plane with the exact triple/ray counts."""
    import multiprocessing as mp
    s_nside = 1 << sample_order
    s_lmax = 2 * s_nside
    s_ray_order = sample_order
    jobs = [(sample_order, s_lmax, s_ray_order, seed + i) for i in range(cores)]
    t0 = time.time()
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_sample_worker, jobs)
    else:
        res = [_cpu_sample_worker(jobs[0])]
    wall = time.time() - t0
    t_ana = float(np.mean([r[0] for r in res])); t_syn = float(np.mean([r[1] for r in res])); t_ray = float(np.mean([r[2] for r in res]))
    tri_s = triple_count(s_nside, s_lmax); tri_t = triple_count(target_nside, target_lmax)
    nrays_s = 12 * (1 << (2 * s_ray_order))
    per_plane_core_s = (t_ana + t_syn) * (tri_t / tri_s) + t_ray * (target_nrays / nrays_s)
    planes_per_s = cores / per_plane_core_s
    sample = ("reference map2alm_mpi+filter+alm2allmaps_mpi+shearinterp_comp+rayprop_sphere (oracle/_ref, single-rank MPI stub, "
              "FFTW replaced by the FP64 shim) at Nside=%d lmax=%d rays Nside=%d on %d process(es) in parallel: "
              "%.2f s analysis, %.2f s synthesis, %.2f s rays per process (wall %.1f s); EXTRAPOLATED to Nside=%d lmax=%d, %d rays "
              "with the exact triple ratio %.1f and ray ratio %.1f" % (
                  s_nside, s_lmax, 1 << s_ray_order, cores, t_ana, t_syn, t_ray, wall, target_nside, target_lmax, target_nrays,
                  tri_t / tri_s, target_nrays / nrays_s))
    return planes_per_s, sample, wall


def host_cores(sample_order):
    """Processes for the CPU legs: every host core, bounded by memory (one reference process holds ~0.2 GB at sample
    order 8, ~0.8 GB at order 9: six maps plus 176-byte rays) so a many-core box is not driven out of RAM."""
    cores = os.cpu_count() or 1
    try:
        import psutil
        per_proc = 0.25e9 * 4 ** max(sample_order - 8, 0)
        cores = max(1, min(cores, int(0.5 * psutil.virtual_memory().available / per_proc)))
    except Exception:
        cores = min(cores, 32)
    return cores


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    cfg = workload_config(a)
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcalclens_ref.so not built (needs /root/reference at build time)"}))
        return 0
    cores = host_cores(a.ref_sample_order)
    vals, wall = [], 0.0
    sample_txt = ""
    for i in range(a.warmup + a.steps):
        v, sample_txt, w = cpu_reference_sample(a.ref_sample_order, cores, a.nside, a.lmax, cfg["nrays"], seed=100 + i)
        if i >= a.warmup:
            vals.append(v); wall += w
    value = float(np.mean(vals)) if vals else 0.0
    line = {"metric": METRIC, "value": value, "unit": "planes/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 / value if value else None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": cfg["config"],
            "ray_plane_updates_per_s": value * cfg["nrays"],
            "cpu_baseline": {"value": value, "unit": "planes/s", "cores": cores, "kind": "reference", "sample": sample_txt},
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


def make_count_maps(solver, nmaps, lmax, seed=1234, nbar=8.0, sigma=0.5):
    """Synthetic lognormal over-density shells: Gaussian field with C_l ~ (l+10)^-1.2 synthesised
    with this library's own alm->map, delta = exp(sigma g - sigma^2/2) - 1, counts = nbar (1 + delta).  Outside any timed region."""
    import torch
    p = solver.plan
    maps = []
    for k in range(nmaps):
        gen = torch.Generator(device=solver.device); gen.manual_seed(seed + k + 7919 * solver.rank)
        are = torch.randn(max(p.Nlm, 1), generator=gen, device=solver.device, dtype=torch.float64)
        aim = torch.randn(max(p.Nlm, 1), generator=gen, device=solver.device, dtype=torch.float64)
        # scale by sqrt(C_l/2): degree l of every local (m, l) without one launch per m
        if p.nm_loc:
            mloc = torch.as_tensor(np.asarray(p.m_local), device=solver.device, dtype=torch.int64)
            cnt = lmax + 1 - mloc
            start = torch.cumsum(cnt, 0) - cnt
            ls = (torch.arange(int(cnt.sum()), device=solver.device, dtype=torch.int64)
                  - torch.repeat_interleave(start, cnt) + torch.repeat_interleave(mloc, cnt)).double()
            ms = torch.repeat_interleave(mloc, cnt)
        else:
            ls = torch.zeros(1, device=solver.device, dtype=torch.float64); ms = torch.zeros(1, device=solver.device, dtype=torch.int64)
        amp = torch.sqrt(0.5 * (ls + 10.0) ** -1.2)
        amp[ls < 1] = 0.0
        are[:ls.numel()] *= amp; aim[:ls.numel()] *= amp
        aim[:ls.numel()][ms == 0] = 0.0          # m = 0 coefficients are real
        solver.alm2allmaps(are, aim)
        g = solver.maps[0].double()
        g = g / g.std()
        counts = (nbar * torch.exp(sigma * g - 0.5 * sigma * sigma)).float()
        maps.append(counts)
    return maps


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
    ap.add_argument("--ref-sample-order", type=int, default=8, help="log2 Nside of the bounded CPU sample of the reference arm")
    ap.add_argument("--cpu-baseline-order", type=int, default=9, help="log2 Nside of the cpu_baseline sample inside the GPU arm (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap-rays", type=int, default=0,
                    help="1: run the ray kernel of plane p on its own stream beside the next plane's Legendre analysis")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="multi-GPU exchange: stores into peer memory from the producing kernels, or NCCL all-to-all + all-reduce")
    a = ap.parse_args()
    if a.lmax is None:
        a.lmax = 2 * a.nside
    if a.ray_nside is None:
        a.ray_nside = a.nside
    if a.impl == "reference":
        return run_reference_arm(a)

    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner, ...) go to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import calclens_b200 as clb
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

    t_setup = time.time()
    solver = poisson.LensPlaneSolver(order, a.lmax, ray_order, dist_group=group, device=local_rank, fused=(a.exchange == "fused"),
                                     overlap_rays=bool(a.overlap_rays))
    cosmo = poisson.Cosmology(0.27)
    max_dist = 30.0 * a.planes
    pp = [poisson.plane_params(p, a.planes, max_dist, 0.27, cosmo) for p in range(a.planes)]
    solver.init_rays(pp[0]["binL"] / 2.0)
    nmaps = 3
    dev_maps = make_count_maps(solver, nmaps, a.lmax)
    # partMass so that the shell holds its mean mass: sum(mass) = Omega_m rho_crit V_shell  
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
        part_mass = 0.27 * poisson.RHO_CRIT * vshell / (8.0 * solver.npix)
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

    # stage events (for the per-kernel roofline), recorded on the launching stream inside the timed region
    stage_names = ["scale", "fft_analysis", "a2a_g", "legendre_analysis", "legendre_synthesis", "a2a_b", "fft_synthesis", "map_allreduce", "rays"]
    stage_ms = {k: 0.0 for k in stage_names}

    def timed_step(step, src_maps, record):
        premul, densmul, backdens, wpp1, wp, wpm1 = plane_args(step)
        p = solver.plan
        ev = []

        def mark(*_):
            if record:
                e = torch.cuda.Event(enable_timing=True); e.record(); ev.append(e)
        mark()
        solver.load_density(src_maps[step % nmaps], premul, densmul, backdens); mark()
        solver.solve(mark=lambda name: mark())
        solver.ray_update(wpp1, wp, wpm1); mark()
        return ev

    # ---- device-resident throughput ("value")
    for s in range(a.warmup):
        timed_step(s, dev_maps, False)
    solver.sync_rays()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.clb_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    all_ev = []
    torch.cuda.profiler.start()    # lets `ncu --profile-from-start off` list exactly the timed region's launches
    e0.record()
    for s in range(a.steps):
        all_ev.append(timed_step(a.warmup + s, dev_maps, True))
    solver.sync_rays()      # (overlap_rays: the last plane's ray kernel is part of the timed region)
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    launches = L.clb_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    for ev in all_ev:
        for k, name in enumerate(stage_names):
            stage_ms[name] += ev[k].elapsed_time(ev[k + 1]) / a.steps
    if solver.overlap_rays and solver.ray_events:
        # the ray kernel ran beside the next plane's Legendre analysis: report its own (stretched) duration
        evs = solver.ray_events[-a.steps:]
        stage_ms["rays"] = sum(x.elapsed_time(y) for x, y in evs) / len(evs)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / a.steps
    value = 1000.0 / ms_per_step

    # ---- end-to-end through the public API with host buffers (H2D of the plane's map + D2H of the summary every step)
    base = a.warmup + a.steps      # planes continue where the first loop stopped (rays sit at that shell)
    for s in range(min(a.warmup, 2)):
        nxt = (host_maps[(base + s + 1) % nmaps],) + tuple(plane_args(base + s + 1, peek=True)[:3])
        solver.step(host_maps[(base + s) % nmaps], *plane_args(base + s), prefetch=nxt)
    base += min(a.warmup, 2)
    barrier()
    e0.record()
    for s in range(a.steps):
        # the next plane's map starts streaming in (pinned host -> this rank's rings on the device) behind this plane's kernels
        nxt = (host_maps[(base + s + 1) % nmaps],) + tuple(plane_args(base + s + 1, peek=True)[:3])
        summ = solver.step(host_maps[(base + s) % nmaps], *plane_args(base + s), prefetch=nxt)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / a.steps
    e2e_value = 1000.0 / e2e_ms

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
        ach_syn = wm["flops_synthesis"] / world / (stage_ms["legendre_synthesis"] * 1e-3) / 1e12
        traffic = None
        try:
            tj = json.load(open(os.path.join(HERE, "profiles", "r01_traffic.json")))
            if world == 1 and a.nside == 4096 and a.lmax == 8192:
                traffic = tj.get("legendre_synthesis", {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": "legendre_synthesis_kernel (dominant)", "bound": "fp64", "achieved": ach_syn, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach_syn / fp64_peak, "traffic": traffic,
                    "peak_source": fp64_src, "algorithmic_flops_per_launch": wm["flops_synthesis"] / world,
                    "note": "FP64 FMA-pipe roofline (not tensor: B200's DMMA rate equals its DFMA rate and the recurrence is serial per ring). "
                            "algorithmic = 16 flop per (m, ring pair, l) triple of the reference's lmin cut: 2-instruction recurrence + three "
                            "complex sums (DESIGN.md section 4); with SURVEY.md 8d's four-sum count of 20 the same time gives %.2f TFLOP/s. "
                            "traffic = dram bytes of one launch from profiles/ (ncu --set full)" % (
                                wm["flops_synthesis_survey"] / world / (stage_ms["legendre_synthesis"] * 1e-3) / 1e12)}
        stages = {
            "legendre_analysis": {"bound": "fp64", "achieved": wm["flops_analysis"] / world / (stage_ms["legendre_analysis"] * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s"},
            "fft_analysis": {"bound": "hbm", "achieved": wm["bytes_fft_analysis"] / world / (stage_ms["fft_analysis"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"},
            "fft_synthesis": {"bound": "hbm", "achieved": wm["bytes_fft_synthesis"] / world / (stage_ms["fft_synthesis"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"},
            "rays": {"bound": "hbm", "achieved": wm["bytes_rays"] / world / (stage_ms["rays"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"},
        }
        for v in stages.values():
            v["frac"] = v["achieved"] / v["peak"]
        cpu_baseline = None
        if world == 1 and not a.no_cpu_baseline and a.cpu_baseline_order > 0:
            try:
                from oracle import ref
                if ref.available():
                    cores = host_cores(a.cpu_baseline_order)
                    v, txt, _ = cpu_reference_sample(a.cpu_baseline_order, cores, a.nside, a.lmax, nrays_total)
                    cpu_baseline = {"value": v, "unit": "planes/s", "cores": cores, "kind": "reference", "sample": txt}
                else:
                    cpu_baseline = {"value": None, "unit": "planes/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
            except Exception as exc:   # the GPU numbers stand even if the CPU leg fails
                cpu_baseline = {"value": None, "unit": "planes/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (exc,)}
        line = {"metric": METRIC, "value": value, "unit": "planes/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg["config"],
                "ray_plane_updates_per_s": value * nrays_total,
                "e2e": {"value": e2e_value, "unit": "planes/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 4 * solver.npix,
                        "d2h_bytes_per_step": 48 * world, "api": "calclens_b200.poisson.LensPlaneSolver.step(pinned host count map, prefetch=next plane) -> 6 ray sums on host; every rank reads only its own rings of the host map",
                        "last_summary": [float(x) for x in summ]},
                "gpu_launches": int(launches * world),
                "clocks": clocks,
                "roofline": roofline, "roofline_stages": stages, "stage_ms": stage_ms,
                "hbm_peak_source": hbm_src,
                "overlap_rays": bool(a.overlap_rays),
                "exchange": ("fused peer stores (CUDA IPC over NVLink) + stream barriers" if solver.fused else
                             "NCCL all-to-all-v + all-reduce" if world > 1 else "none (single GPU)"),
                "cpu_baseline": cpu_baseline,
                "setup_s": t_setup}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    solver.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
