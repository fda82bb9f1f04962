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
                 waves=None, amp=0.3, dens_scale=1.0):
    """Analytic thin torus evaluated at Cartesian KS points -> (dens, velx, vely, velz, eint, b1, b2, b3)."""
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
                            dens_scale=1.0, dtype=np.float64):
    """Single-level cube ``[-extent, extent]^3`` of ``ncells^3`` cells in ``(ncells/block)^3`` meshblocks.

    Returns a dict with the AthenaK arrays plus ``VariableNames`` and ``fluid_gamma``.
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
    x1v = np.empty((nmb, block)); x2v = np.empty((nmb, block)); x3v = np.empty((nmb, block))
    x1f = np.empty((nmb, block + 1)); x2f = np.empty((nmb, block + 1)); x3f = np.empty((nmb, block + 1))
    loc = np.empty((nmb, 3), dtype=np.int64)
    ar = np.arange(block + 1)
    mb = 0
    for lk in range(nb):
        for lj in range(nb):
            for li in range(nb):
                f1 = -extent + (li * block + ar) * dx
                f2 = -extent + (lj * block + ar) * dx
                f3 = -extent + (lk * block + ar) * dx
                x1f[mb], x2f[mb], x3f[mb] = f1, f2, f3
                x1v[mb], x2v[mb], x3v[mb] = f1[:-1] + dx / 2, f2[:-1] + dx / 2, f3[:-1] + dx / 2
                loc[mb] = (li, lj, lk)
                zz, yy, xx = np.meshgrid(x3v[mb], x2v[mb], x1v[mb], indexing='ij')   # [k, j, i]
                fl = torus_fields(xx, yy, zz, fluid_gamma=fluid_gamma, waves=(kvec, phase), amp=amp,
                                  dens_scale=dens_scale)
                for q in range(5):
                    uov[q, mb] = fl[q].astype(np.float32)
                for q in range(3):
                    B[q, mb] = fl[5 + q].astype(np.float32)
                mb += 1
    return dict(uov=uov, B=B, x1v=x1v, x2v=x2v, x3v=x3v, x1f=x1f, x2f=x2f, x3f=x3f,
                LogicalLocations=loc, Levels=np.zeros(nmb, dtype=np.int64),
                VariableNames=VARIABLE_NAMES, fluid_gamma=fluid_gamma)
