"""TEST INFRASTRUCTURE (oracle): literal restatement of the reference loader's ghost-zone fill across refinement
levels, /root/reference/mahakala/grmhd/athenak.py:160-514 (`_get_key_for_level`, `_get_new_meshblock_boundary`).

Only ``tests/`` may import this module.  It follows the reference direction by direction:

* same-level neighbour (athenak.py:213-229): copy of the neighbour's edge layer;
* coarser neighbour (:231-295): the half of the coarse block selected by the parity of the target location is
  injected into every other fine ghost cell (two fine cells per coarse cell and axis);
* finer neighbours (:304-512): each ghost cell accumulates ``contribution / 8`` of the 8 fine cells it covers, in
  the loop order v1 (i offset) outermost ... v3 (k offset) innermost, starting from zero.

The reference's EDGE branch of the finer-neighbour case (:347-424) assigns its stride slices to ``source_*`` instead
of ``copy_source_*``, so its edge ghost cells are wrong; ``fill_direction`` therefore returns
``None`` for that case and the tests pin faces and corners only (where the reference is right) against it, and the
edges against the brute-force expectation.
"""
import numpy as np


def key_for_level(c_level, n_level, t, d):
    """athenak.py:160-190: logical location of the neighbour block in direction d, one level up or down."""
    if c_level == n_level:
        return (c_level,) + tuple(t[a] + d[a] for a in range(3))
    if n_level == c_level + 1:
        return (n_level,) + tuple(2 * (t[a] + d[a]) for a in range(3))
    if n_level == c_level - 1:
        return (n_level,) + tuple((t[a] + d[a]) // 2 for a in range(3))
    raise Exception("Unable to compute key for meshblock level")


def fill_direction(data, index, mb, lev, t, d):
    """Ghost cells of block ``mb`` (level lev, logical location t = (ti, tj, tk)) in direction d = (di, dj, dk).

    data: (8, nmb, nk, nj, ni) interior values (uov and B concatenated, file order); index: {(level, li, lj, lk): mb}.
    Returns (target index tuple (k, j, i) into the padded block, values (8, ...)); zeros beyond the domain; None for
    the finer-neighbour edge branch, which is buggy in the reference.
    """
    n = (data.shape[4], data.shape[3], data.shape[2])                # (ni, nj, nk)
    tgt = [(-1 if d[a] == 1 else 0) if d[a] else slice(1, n[a] + 1) for a in range(3)]
    # ---- same level (athenak.py:213-229) ----
    nb = index.get((lev,) + tuple(t[a] + d[a] for a in range(3)))
    if nb is not None:
        src = [(0 if d[a] == 1 else -1) if d[a] else slice(0, n[a]) for a in range(3)]
        return (tgt[2], tgt[1], tgt[0]), data[:, nb, src[2], src[1], src[0]]
    # ---- one level up: coarser neighbour, injection (athenak.py:231-295) ----
    nb = index.get(key_for_level(lev, lev - 1, t, d))
    if nb is not None:
        odd = [(t[a] + d[a]) % 2 for a in range(3)]
        out = np.zeros((8,) + tuple((1 if d[a] else n[a]) for a in (2, 1, 0)))
        src = []
        for a in range(3):
            if d[a] == 1:
                src.append(n[a] // 2 if odd[a] else 0)                 # _get_01_source(0, ...)
            elif d[a] == -1:
                src.append(n[a] - 1 if odd[a] else n[a] // 2 - 1)      # _get_01_source(-1, ...)
            else:
                src.append(slice(n[a] // 2, n[a]) if odd[a] else slice(0, n[a] // 2))
        block = data[:, nb][(slice(None),) + tuple(src[a] if not isinstance(src[a], int) else slice(src[a], src[a] + 1)
                                                   for a in (2, 1, 0))]
        for s_i in range(2):
            for s_j in range(2):
                for s_k in range(2):
                    sel = [slice(None)]
                    for a, s_ in ((2, s_k), (1, s_j), (0, s_i)):
                        sel.append(slice(0, 1) if d[a] else slice(s_, n[a] + s_, 2))
                    out[tuple(sel)] = block
        return (tgt[2], tgt[1], tgt[0]), out.reshape((8,) + tuple(m for a, m in zip((2, 1, 0), out.shape[1:]) if not d[a]))
    # ---- one level down: finer neighbours, mean of 8 (athenak.py:304-512) ----
    num_slices = sum(1 for a in range(3) if d[a] == 0)
    base = key_for_level(lev, lev + 1, t, d)
    if base not in index:                                            # beyond the domain: the ghost cells stay zero
        shape = (8,) + tuple(n[a] for a in (2, 1, 0) if not d[a])
        return (tgt[2], tgt[1], tgt[0]), np.zeros(shape)
    if num_slices == 1:
        return None                                                  # reference's edge branch is buggy (2.2 #9)
    new = list(base[1:])
    src0 = [0, 0, 0]
    for a in range(3):
        if d[a] == -1:
            new[a] += 1
            src0[a] = n[a] - 2
    free = [a for a in range(3) if d[a] == 0]
    out = np.zeros((8,) + tuple((1 if d[a] else n[a]) for a in (2, 1, 0)))
    pos_ranges = [range(2) if a in free else range(1) for a in range(3)]
    for p0 in pos_ranges[0]:
        for p1 in pos_ranges[1]:
            for p2 in pos_ranges[2]:
                pos = (p0, p1, p2)
                fm = index.get((lev + 1, new[0] + pos[0], new[1] + pos[1], new[2] + pos[2]))
                if fm is None:
                    return None
                tsel = [slice(None)]
                for a in (2, 1, 0):
                    tsel.append(slice(0, 1) if d[a] else slice(pos[a] * n[a] // 2, (1 + pos[a]) * n[a] // 2))
                for v1 in range(2):
                    for v2 in range(2):
                        for v3 in range(2):
                            v = (v1, v2, v3)
                            ssel = [slice(None)]
                            for a in (2, 1, 0):
                                ssel.append(slice(src0[a] + v[a], src0[a] + v[a] + 1) if d[a] else slice(v[a], n[a] + v[a], 2))
                            out[tuple(tsel)] += data[:, fm][tuple(ssel)] / 8.
    return (tgt[2], tgt[1], tgt[0]), out.reshape((8,) + tuple(m for a, m in zip((2, 1, 0), out.shape[1:]) if not d[a]))
