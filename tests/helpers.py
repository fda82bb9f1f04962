"""Shared builders for the fluid-model tests (oracle model and device model from the same arrays)."""
import numpy as np

from mahakala_b200.synthetic import make_synthetic_snapshot

M_BH = 6.2e9 * 1.989e33
MASS_SCALE = 1.e26


def snapshot_arrays(ncells=32, block=16, extent=16.0, seed=0, **kw):
    return make_synthetic_snapshot(ncells=ncells, block=block, extent=extent, seed=seed, **kw)


def oracle_model(arr, bhspin):
    from oracle import mahakala_oracle as onp
    return onp.AthenakFluidModel(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                 arr["x3f"], arr["LogicalLocations"], arr["Levels"], bhspin,
                                 fluid_gamma=arr["fluid_gamma"], variable_names=arr["VariableNames"])


def device_model(arr, bhspin, **kw):
    from mahakala_b200.grmhd import AthenakFluidModel
    return AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"],
                                         arr["x2f"], arr["x3f"], arr["LogicalLocations"], arr["Levels"], bhspin,
                                         fluid_gamma=arr["fluid_gamma"], VariableNames=arr["VariableNames"], **kw)


def close(a, b, rtol, atol):
    a = np.asarray(a); b = np.asarray(b)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))


def two_level_mesh(n=8, seed=0, fluid_gamma=13. / 9):
    """A two-level AthenaK-style mesh: 2x2x2 root blocks of n^3 cells on [-8, 8]^3, the root block at logical
    location (1, 1, 1) replaced by its 8 children (level 1).  Fields come from one global fine-resolution
    array F (level-1 cells) and its 2x2x2 restriction C (level-0 cells), so that the expected ghost cells of
    every block follow from plain indexing of F and C (brute-force oracle for the AMR ghost fill).

    Returns (arrays dict for from_arrays, expected all_meshblocks).
    """
    rng = np.random.default_rng(seed)
    nf = 4 * n                       # fine cells per side over the whole domain
    F = np.empty((8, nf, nf, nf))
    zz, yy, xx = np.meshgrid(*(3 * [(-8 + (np.arange(nf) + 0.5) * 16.0 / nf)]), indexing='ij')
    from mahakala_b200.synthetic import torus_fields
    fl = torus_fields(xx, yy, zz, fluid_gamma=fluid_gamma, R0=4.0, R_in=1.0)
    order = [0, 1, 2, 3, 4, 5, 6, 7]          # dens, velx, vely, velz, eint, b1, b2, b3 (uov then B)
    for q in range(8):
        F[q] = fl[order[q]] * (1.0 + 0.05 * rng.standard_normal((nf, nf, nf)))
    F = F.astype(np.float32).astype(np.float64)
    C = F.reshape(8, nf // 2, 2, nf // 2, 2, nf // 2, 2).mean(axis=(2, 4, 6))
    blocks = []                      # (level, li, lj, lk)
    for lk in range(2):
        for lj in range(2):
            for li in range(2):
                if (li, lj, lk) != (1, 1, 1):
                    blocks.append((0, li, lj, lk))
    for lk in range(2, 4):
        for lj in range(2, 4):
            for li in range(2, 4):
                blocks.append((1, li, lj, lk))
    nmb = len(blocks)
    uov = np.empty((5, nmb, n, n, n)); B = np.empty((3, nmb, n, n, n))
    xv = [np.empty((nmb, n)) for _ in range(3)]
    xf = [np.empty((nmb, n + 1)) for _ in range(3)]
    expected = np.zeros((nmb, 8, n + 2, n + 2, n + 2))

    def sample_global(level, gk, gj, gi):
        """value of the mesh at global cell (gk, gj, gi) of `level`, as a block of that level would see it"""
        size = 2 * n * (2 ** level)
        if not (0 <= gk < size and 0 <= gj < size and 0 <= gi < size):
            return None
        if level == 0:
            return C[:, gk, gj, gi]          # refined region: average of the 8 children = C by construction
        refined = gk >= nf // 2 and gj >= nf // 2 and gi >= nf // 2
        return F[:, gk, gj, gi] if refined else C[:, gk // 2, gj // 2, gi // 2]

    for mb, (lev, li, lj, lk) in enumerate(blocks):
        dx = 16.0 / (2 * n * 2 ** lev)
        src = C if lev == 0 else F
        for ax, l in enumerate((li, lj, lk)):
            f = -8 + (l * n + np.arange(n + 1)) * dx
            xf[ax][mb] = f
            xv[ax][mb] = f[:-1] + dx / 2
        blk = src[:, lk * n:(lk + 1) * n, lj * n:(lj + 1) * n, li * n:(li + 1) * n]
        uov[:, mb] = blk[:5]
        B[:, mb] = blk[5:]
        for k in range(n + 2):
            for j in range(n + 2):
                for i in range(n + 2):
                    v = sample_global(lev, lk * n + k - 1, lj * n + j - 1, li * n + i - 1)
                    if v is not None:
                        expected[mb, :, k, j, i] = v
    arr = dict(uov=uov, B=B, x1v=xv[0], x2v=xv[1], x3v=xv[2], x1f=xf[0], x2f=xf[1], x3f=xf[2],
               LogicalLocations=np.array([[b[1], b[2], b[3]] for b in blocks]), Levels=np.array([b[0] for b in blocks]),
               VariableNames=('dens', 'velx', 'vely', 'velz', 'eint', 'bcc1', 'bcc2', 'bcc3'), fluid_gamma=fluid_gamma)
    return arr, expected


def noncubic_mesh(nb=(2, 3, 2), n=(8, 6, 4), seed=3, fluid_gamma=13. / 9):
    """Single-level mesh with NON-cubic meshblocks: nb = blocks along (x1, x2, x3), n = cells per block along
    (x1, x2, x3) = (ni, nj, nk), cell size 0.5 on every axis.  Returns (arrays for from_arrays, global zero-padded
    field array G (8, N3+2, N2+2, N1+2), global cell-centre coordinates) so that ghost zones and trilinear samples
    can be checked by plain indexing of G."""
    rng = np.random.default_rng(seed)
    ni, nj, nk = n
    N1, N2, N3 = nb[0] * ni, nb[1] * nj, nb[2] * nk
    dx = 0.5
    lo = (-N1 * dx / 2, -N2 * dx / 2, -N3 * dx / 2)
    G = np.zeros((8, N3 + 2, N2 + 2, N1 + 2))
    G[:, 1:-1, 1:-1, 1:-1] = rng.uniform(0.5, 1.5, (8, N3, N2, N1)).astype(np.float32)
    nmb = nb[0] * nb[1] * nb[2]
    uov = np.empty((5, nmb, nk, nj, ni)); B = np.empty((3, nmb, nk, nj, ni))
    xv = [np.empty((nmb, m)) for m in n]
    xf = [np.empty((nmb, m + 1)) for m in n]
    loc = np.empty((nmb, 3), dtype=np.int64)
    mb = 0
    for lk in range(nb[2]):
        for lj in range(nb[1]):
            for li in range(nb[0]):
                blk = G[:, 1 + lk * nk:1 + (lk + 1) * nk, 1 + lj * nj:1 + (lj + 1) * nj, 1 + li * ni:1 + (li + 1) * ni]
                uov[:, mb] = blk[:5]
                B[:, mb] = blk[5:]
                for ax, (l, m) in enumerate(zip((li, lj, lk), n)):
                    f = lo[ax] + (l * m + np.arange(m + 1)) * dx
                    xf[ax][mb] = f
                    xv[ax][mb] = f[:-1] + dx / 2
                loc[mb] = (li, lj, lk)
                mb += 1
    arr = dict(uov=uov, B=B, x1v=xv[0], x2v=xv[1], x3v=xv[2], x1f=xf[0], x2f=xf[1], x3f=xf[2], LogicalLocations=loc,
               Levels=np.zeros(nmb, dtype=np.int64),
               VariableNames=('dens', 'velx', 'vely', 'velz', 'eint', 'bcc1', 'bcc2', 'bcc3'), fluid_gamma=fluid_gamma)
    centres = [lo[ax] + (np.arange(-1, (N1, N2, N3)[ax] + 1) + 0.5) * dx for ax in range(3)]      # incl. the zero padding
    return arr, G, centres


def three_level_mesh(n=4, seed=1, fluid_gamma=13. / 9):
    """Three refinement levels with 2:1 balance: 2x2x2 root blocks of n^3 cells on [-8, 8]^3; root block (1,1,1) is
    replaced by its 8 children (level 1) and the level-1 child in the domain corner, (3,3,3), by its 8 children
    (level 2): 7 + 7 + 8 = 22 meshblocks.  Level-l data are the 2x2x2 restriction of the level-(l+1) data, so every
    ghost cell has a brute-force expectation: same-level / coarser leaf -> that level's array (injection), finer
    leaves -> the restriction, outside the domain -> 0.  Returns (arrays for from_arrays, expected all_meshblocks)."""
    rng = np.random.default_rng(seed)
    L = 2
    nf = 2 * n * 2 ** L
    A = {L: rng.uniform(0.5, 1.5, (8, nf, nf, nf)).astype(np.float32).astype(np.float64)}
    for lev in range(L - 1, -1, -1):
        f = A[lev + 1]
        m = f.shape[1] // 2
        A[lev] = f.reshape(8, m, 2, m, 2, m, 2).mean(axis=(2, 4, 6))
    leaves = [(0, li, lj, lk) for lk in range(2) for lj in range(2) for li in range(2) if (li, lj, lk) != (1, 1, 1)]
    leaves += [(1, li, lj, lk) for lk in (2, 3) for lj in (2, 3) for li in (2, 3) if (li, lj, lk) != (3, 3, 3)]
    leaves += [(2, li, lj, lk) for lk in (6, 7) for lj in (6, 7) for li in (6, 7)]
    leafset = set(leaves)

    def value_at(level, gk, gj, gi):
        size = 2 * n * 2 ** level
        if not (0 <= gk < size and 0 <= gj < size and 0 <= gi < size):
            return None
        for lev2 in range(level, -1, -1):                      # covered by a leaf of this or a coarser level
            sh = level - lev2
            ck, cj, ci = gk >> sh, gj >> sh, gi >> sh
            if (lev2, ci // n, cj // n, ck // n) in leafset:
                return A[lev2][:, ck, cj, ci]
        return A[level][:, gk, gj, gi]                          # covered by finer leaves: their restriction

    nmb = len(leaves)
    uov = np.empty((5, nmb, n, n, n)); B = np.empty((3, nmb, n, n, n))
    xv = [np.empty((nmb, n)) for _ in range(3)]
    xf = [np.empty((nmb, n + 1)) for _ in range(3)]
    expected = np.zeros((nmb, 8, n + 2, n + 2, n + 2))
    for mb, (lev, li, lj, lk) in enumerate(leaves):
        dx = 16.0 / (2 * n * 2 ** lev)
        for ax, l in enumerate((li, lj, lk)):
            f = -8 + (l * n + np.arange(n + 1)) * dx
            xf[ax][mb] = f
            xv[ax][mb] = f[:-1] + dx / 2
        blk = A[lev][:, lk * n:(lk + 1) * n, lj * n:(lj + 1) * n, li * n:(li + 1) * n]
        uov[:, mb] = blk[:5]
        B[:, mb] = blk[5:]
        for k in range(n + 2):
            for j in range(n + 2):
                for i in range(n + 2):
                    v = value_at(lev, lk * n + k - 1, lj * n + j - 1, li * n + i - 1)
                    if v is not None:
                        expected[mb, :, k, j, i] = v
    arr = dict(uov=uov, B=B, x1v=xv[0], x2v=xv[1], x3v=xv[2], x1f=xf[0], x2f=xf[1], x3f=xf[2],
               LogicalLocations=np.array([[b[1], b[2], b[3]] for b in leaves]), Levels=np.array([b[0] for b in leaves]),
               VariableNames=('dens', 'velx', 'vely', 'velz', 'eint', 'bcc1', 'bcc2', 'bcc3'), fluid_gamma=fluid_gamma)
    return arr, expected
