"""Generate tests/golden/*.npz from the compiled reference (oracle/_ref).  Run here, where /root/reference exists:
    python tools/make_golden.py
The vectors are small on purpose; every array is produced by the UNMODIFIED reference functions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
rng = np.random.default_rng(20261018)
R = ref.lib()

# ---- ring weights (HEALPix data shipped with the reference, healpix_weights/weight_ring_n*.fits)
wts = {}
for order in range(1, 11):
    wts["n%05d" % (1 << order)] = ref.read_ring_weights("/root/reference/healpix_weights", order)
np.savez_compressed(os.path.join(OUT, "ring_weights.npz"), **wts)

# ---- indexing
idx = {}
for order in (0, 1, 2, 3, 4, 5):
    npix = 12 << (2 * order)
    p = np.arange(npix)
    idx["ring2nest_o%d" % order] = np.array([R.ring2nest(int(i), order) for i in p])
    idx["nest2ring_o%d" % order] = np.array([R.nest2ring(int(i), order) for i in p])
    idx["nest2peano_o%d" % order] = np.array([R.nest2peano(int(i), order) for i in p])
th = np.arccos(rng.uniform(-1, 1, 3000)); ph = rng.uniform(0, 2 * np.pi, 3000)
# edge cases: poles, equator, face boundaries, phi wrap
th[:8] = [0.0, np.pi, np.pi / 2, np.arccos(2 / 3), np.arccos(-2 / 3), 1e-9, np.pi - 1e-9, np.pi / 2]
ph[:8] = [0.0, 0.0, 0.0, np.pi / 2, np.pi, 2 * np.pi - 1e-12, 1e-12, np.pi / 4]
idx["theta"] = th; idx["phi"] = ph
for order in (0, 3, 8, 13):
    idx["ang2nest_o%d" % order] = np.array([R.ang2nest(float(t), float(q), order) for t, q in zip(th, ph)])
for order in (1, 4, 10):
    pix = np.zeros((th.size, 4), dtype=np.int64); wgt = np.zeros((th.size, 4))
    for i, (t, q) in enumerate(zip(th, ph)):
        if t == 0.0 or t == np.pi:
            t = 1e-7 if t == 0.0 else np.pi - 1e-7
        a, b = ref.get_interpol(t, q, order)
        pix[i] = a; wgt[i] = b
    idx["interpol_pix_o%d" % order] = pix; idx["interpol_wgt_o%d" % order] = wgt
np.savez_compressed(os.path.join(OUT, "healpix_index.npz"), **idx)

# ---- SHT
sht = {}
for tag, order, lmax, useW in (("a", 3, 23, True), ("b", 3, 16, False), ("c", 4, 32, True), ("d", 1, 5, False)):
    nside = 1 << order; npix = 12 * nside * nside
    m = (rng.lognormal(sigma=0.5, size=npix) * 8.0).astype(np.float32)
    m = (m * np.float32(3e-4) - np.float32(8 * np.exp(0.125) * 3e-4)).astype(np.float32)
    w = wts["n%05d" % nside] if useW else None
    are, aim = ref.map2alm(order, lmax, m, w)
    fre, fim = ref.poisson_filter(lmax, are, aim)
    maps = ref.alm2allmaps(order, lmax, fre, fim)
    sht[tag + "_order"] = order; sht[tag + "_lmax"] = lmax; sht[tag + "_weights"] = int(useW)
    sht[tag + "_map"] = m; sht[tag + "_alm_re"] = are; sht[tag + "_alm_im"] = aim
    sht[tag + "_falm_re"] = fre; sht[tag + "_falm_im"] = fim; sht[tag + "_maps"] = maps
# point mass (single non-zero RING pixel), Nside 16, lmax 47
order, lmax = 4, 47
npix = 12 << (2 * order)
pm = np.zeros(npix, dtype=np.float32); pm[1000] = 1.0
are, aim = ref.map2alm(order, lmax, pm)
fre, fim = ref.poisson_filter(lmax, are, aim)
sht["pm_pixel"] = 1000; sht["pm_maps"] = ref.alm2allmaps(order, lmax, fre, fim)
# lambda_lm spot values
vals = []
for (lmax_, cth, m_) in ((64, 0.3, 0), (64, 0.3, 5), (64, -0.7, 17), (200, 0.999, 30), (200, 0.05, 150)):
    firstl, vec = ref.plmgen(lmax_, cth, np.sqrt((1 - cth) * (1 + cth)), m_)
    vals.append((lmax_, cth, m_, firstl, vec[firstl:firstl + 8].copy()))
sht["plm_args"] = np.array([(v[0], v[1], v[2], v[3]) for v in vals])
sht["plm_vals"] = np.array([v[4] for v in vals])
np.savez_compressed(os.path.join(OUT, "sht.npz"), **sht)

# ---- rays
order = 4
npix = 12 << (2 * order)
maps = (rng.normal(size=(6, npix)) * np.array([1, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2])[:, None]).astype(np.float32)
rays = ref.init_rays(5, 15.0)[::6].copy()   # 2048 rays spread over the sphere
rays["n"] += rng.normal(size=rays["n"].shape) * 0.02
r0 = rays.copy()
ref.shearinterp(order, 2, maps, rays); r1 = rays.copy()
ref.rayprop(rays, 45.0, 15.0, 0.0); r2 = rays.copy()
rays["alpha"] = 0; rays["U"] = 0; rays["phi"] = 0
ref.shearinterp(order, 2, maps, rays)
ref.rayprop(rays, 75.0, 45.0, 15.0); r3 = rays.copy()
# a ray with zero deflection takes the alpha == 0 branch of rayprop_sphere (rayprop.c:124-133)
rz = ref.init_rays(2, 15.0); rz["U"] = rng.normal(size=rz["U"].shape) * 1e-2
rz0 = rz.copy(); ref.rayprop(rz, 45.0, 15.0, 0.0)
np.savez_compressed(os.path.join(OUT, "rays.npz"), order=order, maps=maps, rays0=r0.view(np.uint8), rays_interp=r1.view(np.uint8),
                    rays_prop1=r2.view(np.uint8), rays_prop2=r3.view(np.uint8), rz0=rz0.view(np.uint8), rz1=rz.view(np.uint8))

# ---- the callers either side of the path ("next" rows): NGP deposit, write_rays' output transform, Born step
nx = {}
order = 3
pos = (rng.normal(size=(5000, 3)) * 200.0).astype(np.float32)
pos[:6] = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [1e-3, -1e-3, 1]]
mass = np.full(5000, 3.7e9, dtype=np.float32)
nx["dep_order"] = order; nx["dep_pos"] = pos; nx["dep_mass"] = mass; nx["dep_map"] = ref.deposit_ngp(pos, mass, order)
ro = ref.init_rays(3, 15.0)
ro["n"] += rng.normal(size=ro["n"].shape) * 0.02
for f, n in (("A", 4), ("Aprev", 4), ("U", 4), ("alpha", 2)):
    ro[f] += rng.normal(size=(ro.size, n)) * 0.05
nx["rays_in"] = ro.copy().view(np.uint8)
r_out = ro.copy(); ref.ray_output(r_out, 3); nx["rays_output"] = r_out.view(np.uint8)
r_born = ro.copy(); ref.rayprop_born(r_born, 45.0, 15.0, 0.0); ref.rayprop_born(r_born, 75.0, 45.0, 15.0)
nx["rays_born"] = r_born.view(np.uint8)
np.savez_compressed(os.path.join(OUT, "next_rows.npz"), **nx)
print("golden vectors written to", OUT, {f: os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)})
