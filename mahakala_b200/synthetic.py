"""Synthetic AthenaK-shaped GRMHD snapshots (host-side data generation; not on the hot path).

There is no network for real ``.athdf`` files, so tests and ``bench.py`` use smooth random fields laid
out exactly like the arrays the reference reads from an AthenaK dump (``grmhd/athenak.py:79-103``):
``uov (5, nmb, nk, nj, ni)`` = dens, velx, vely, velz, eint; ``B (3, nmb, nk, nj, ni)`` = bcc1..3;
``x{1,2,3}v (nmb, n)``; ``x{1,2,3}f (nmb, n+1)``; ``LogicalLocations (nmb, 3)``; ``Levels (nmb,)``.

Recipe: a Keplerian thin torus (power-law density, toroidal field at fixed
plasma beta) modulated by ``exp(amp * G)``, G = a sum of 8 random plane waves drawn from
``numpy.random.default_rng(seed)``.  Values are rounded through float32 (AthenaK writes float32).  The
cell size is a power of two so that cell/face coordinates are exact in float64.
"""
import numpy as np

VARIABLE_NAMES = ('dens', 'velx', 'vely', 'velz', 'eint', 'bcc1', 'bcc2', 'bcc3')


def torus_fields(x, y, z, fluid_gamma=13. / 9, R0=8.0, R_in=2.5, p=1.5, h=0.3, u0=0.25, beta0=3.0,
                 waves=None, amp=0.3, dens_scale=1.0, funnel=None):
    """Analytic thin torus evaluated at Cartesian KS points -> (dens, velx, vely, velz, eint, b1, b2, b3).

    ``funnel`` (dict or None) adds a magnetised polar funnel, the region real GRMHD snapshots have and the smooth
    torus lacks: inside the cone |z| > slope * R a tenuous, hot, vertically magnetised plasma whose magnetisation
    sigma = b^2 / dens rises from 0 at the funnel wall to ``sigma0`` on the axis, so that rays cross the
    sigma = 100 cut of images.py:116-118 on both sides.  Keys: slope (1.2), width (0.35), dens0 (1e-2), theta_u
    (u / dens, 0.5), sigma0 (400), index (1.5)."""
    R2 = x * x + y * y
    R = np.sqrt(R2) + 1e-12
    r = np.sqrt(R2 + z * z) + 1e-12
    H = h * R
    with np.errstate(over='ignore', under='ignore'):
        taper = np.exp(-(R_in / R)**4)
        dens = dens_scale * (R / R0)**(-p) * np.exp(-z * z / (2. * H * H)) * taper
    eint = u0 * dens * (R0 / r)
    vphi = 0.5 / np.sqrt(1. + R)                     # sub-luminal, Kepler-like falloff
    velx = -vphi * y / R
    vely = vphi * x / R
    velz = 0.02 * z / (1. + r)
    bmag = np.sqrt(2. * eint * (fluid_gamma - 1.) / beta0)
    b1 = -bmag * y / R
    b2 = bmag * x / R
    b3 = 0.1 * bmag
    if funnel is not None:
        fp = dict(slope=1.2, width=0.35, dens0=1e-2, theta_u=0.5, sigma0=400.0, index=1.5)
        fp.update(funnel)
        q = np.clip((np.abs(z) - fp["slope"] * R) / (fp["width"] * r), 0., 1.)
        w = q * q * (3. - 2. * q)                       # smoothstep: 0 at the wall, 1 well inside the cone
        dens_f = fp["dens0"] * (1. + r)**(-fp["index"])
        dens = dens + dens_f * w
        eint = eint + fp["theta_u"] * dens_f * w
        b3 = b3 + np.sqrt(fp["sigma0"] * dens_f) * w * np.where(z >= 0, 1., -1.)
    if waves is not None:
        kvec, phase = waves
        G1 = np.zeros_like(x)
        G2 = np.zeros_like(x)
        for m in range(kvec.shape[0]):
            arg = kvec[m, 0] * x + kvec[m, 1] * y + kvec[m, 2] * z + phase[m]
            G1 += np.sin(arg)
            G2 += np.cos(1.7 * arg)
        G1 /= np.sqrt(kvec.shape[0])
        G2 /= np.sqrt(kvec.shape[0])
        m1 = np.exp(amp * G1)
        m2 = np.exp(0.5 * amp * G2)
        dens = dens * m1
        eint = eint * m1
        b1, b2, b3 = b1 * m2, b2 * m2, b3 * m2
        velx = velx * (1. + 0.1 * G2)
        vely = vely * (1. + 0.1 * G1)
    return dens, velx, vely, velz, eint, b1, b2, b3


def make_synthetic_snapshot(ncells=64, block=16, extent=32.0, seed=0, fluid_gamma=13. / 9, amp=0.3,
                            dens_scale=1.0, dtype=np.float64, funnel=None):
    """Single-level cube ``[-extent, extent]^3`` of ``ncells^3`` cells in ``(ncells/block)^3`` meshblocks.

    Returns a dict with the AthenaK arrays plus ``VariableNames`` and ``fluid_gamma``.  Fields are evaluated
    slab by slab on the global grid (one z-layer of blocks at a time) and cut into blocks, which gives the
    same numbers as a block-by-block evaluation (every cell is a function of its own coordinates only).
    """
    assert ncells % block == 0
    nb = ncells // block
    nmb = nb**3
    dx = 2. * extent / ncells
    rng = np.random.default_rng(seed)
    kvec = rng.normal(0., 2. * np.pi / (2. * extent) * 3., size=(8, 3))
    phase = rng.uniform(0., 2. * np.pi, size=8)
    uov = np.empty((5, nmb, block, block, block), dtype=dtype)
    B = np.empty((3, nmb, block, block, block), dtype=dtype)
    ar = np.arange(block + 1)
    faces = [-extent + (l * block + ar) * dx for l in range(nb)]            # per logical index along an axis
    centres = [f[:-1] + dx / 2 for f in faces]
    gx = np.concatenate(centres)                                            # global cell centres along one axis
    loc = np.empty((nmb, 3), dtype=np.int64)
    x1v = np.empty((nmb, block)); x2v = np.empty((nmb, block)); x3v = np.empty((nmb, block))
    x1f = np.empty((nmb, block + 1)); x2f = np.empty((nmb, block + 1)); x3f = np.empty((nmb, block + 1))
    def slab(lk, lj):
        zz, yy, xx = np.meshgrid(centres[lk], centres[lj], gx, indexing='ij')      # [k, j, I] pencil of blocks
        fl = torus_fields(xx, yy, zz, fluid_gamma=fluid_gamma, waves=(kvec, phase), amp=amp, dens_scale=dens_scale,
                          funnel=funnel)
        m0 = (lk * nb + lj) * nb
        for q in range(8):
            # (k, j, li, i) -> (li, k, j, i)
            v = fl[q].astype(np.float32).reshape(block, block, nb, block).transpose(2, 0, 1, 3)
            dst = uov[q] if q < 5 else B[q - 5]
            dst[m0:m0 + nb] = v

    jobs = [(lk, lj) for lk in range(nb) for lj in range(nb)]
    if len(jobs) >= 16:                 # NumPy ufuncs release the GIL: evaluate pencils on all host cores
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as pool:
            list(pool.map(lambda a: slab(*a), jobs))
    else:
        for a in jobs:
            slab(*a)
    for lk in range(nb):
        for lj in range(nb):
            for li in range(nb):
                mb = (lk * nb + lj) * nb + li
                x1f[mb], x2f[mb], x3f[mb] = faces[li], faces[lj], faces[lk]
                x1v[mb], x2v[mb], x3v[mb] = centres[li], centres[lj], centres[lk]
                loc[mb] = (li, lj, lk)
    return dict(uov=uov, B=B, x1v=x1v, x2v=x2v, x3v=x3v, x1f=x1f, x2f=x2f, x3f=x3f,
                LogicalLocations=loc, Levels=np.zeros(nmb, dtype=np.int64),
                VariableNames=VARIABLE_NAMES, fluid_gamma=fluid_gamma)
