"""AthenaK fluid model on a device-resident, repacked snapshot.

Reference: /root/reference/mahakala/grmhd/athenak.py (AthenakFluidModel, :48-812).  What changes:

* the ghost-padded snapshot ``all_meshblocks (nmb, 8, nk+2, nj+2, ni+2)`` is repacked ONCE into a
  cell-major structure-of-arrays layout in HBM (``cells[mb][k][j][i][8]``, optionally float32 when that is
  lossless) instead of being re-uploaded on every call (athenak.py:693);
* the O(nmb) Python mask loop that assigns points to meshblocks (athenak.py:663-670, the reference's
  dominant wall-time cost) becomes an O(1) lookup in a block grid, verified against the exact face extents
  so that the left-open/right-closed membership rule is reproduced bit for bit;
* sampling, the fluid-frame algebra and (in ``images.make_image``) the transfer solve run in CUDA kernels.

``AthenakFluidModel(filename, bhspin, fluid_gamma)`` keeps the reference signature (h5py if present, else the
built-in minimal HDF5 reader ``_hdf5_min``); ``AthenakFluidModel.from_arrays(...)`` takes the arrays an ``.athdf``
file holds.
"""
import ctypes
import time

import numpy as np

from .. import _cabi
from .._device import DeviceArray, as_device, empty, require_gpu, stream_ptr
from .grmhd import GRMHDFluidModel

def vec_metric(X, bhspin):
    """athenak.py:38-40: covariant metric at a batch of points (the device kernel is batched already)."""
    from ..geodesics import metric
    return metric(X, bhspin)


def vec_imetric(X, bhspin):
    """athenak.py:43-45: contravariant metric at a batch of points."""
    from ..geodesics import imetric
    return imetric(X, bhspin)


CANONICAL_PRIMS = ('dens', 'eint', 'velx', 'vely', 'velz', 'bcc1', 'bcc2', 'bcc3')


def _face_or_interior(d, n):
    """(source index/slice in the neighbour, target index/slice in the padded block) along one axis."""
    if d == 1:
        return 0, n + 1
    if d == -1:
        return n - 1, 0
    return slice(0, n), slice(1, n + 1)


def fill_ghost_zones(uov, B, LogicalLocations, Levels):
    """Build ``all_meshblocks (nmb, 8, nk+2, nj+2, ni+2)`` with one layer of ghost cells.

    Same-level neighbours are copied (athenak.py:208-229).  Ghost cells with no same-level neighbour
    (domain boundary, or a refinement boundary) stay zero / are filled by ``_fill_from_other_levels``.
    Host-side, one-time, like the reference's loader.
    """
    uov = np.asarray(uov, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    nprim, nmb, nk, nj, ni = uov.shape
    index = {}
    for mb in range(nmb):
        li, lj, lk = (int(q) for q in LogicalLocations[mb])
        index[(int(Levels[mb]), li, lj, lk)] = mb
    fast = _fill_single_level_box(uov, B, np.asarray(LogicalLocations), np.asarray(Levels), index)
    if fast is not None:
        return fast, index
    out = np.zeros((nmb, 8, nk + 2, nj + 2, ni + 2))
    out[:, :nprim, 1:-1, 1:-1, 1:-1] = np.moveaxis(uov, 0, 1)
    out[:, nprim:, 1:-1, 1:-1, 1:-1] = np.moveaxis(B, 0, 1)
    multilevel = len(set(int(l) for l in Levels)) > 1
    for (lev, li, lj, lk), mb in index.items():
        for dk in (-1, 0, 1):
            sk, tk = _face_or_interior(dk, nk)
            for dj in (-1, 0, 1):
                sj, tj = _face_or_interior(dj, nj)
                for di in (-1, 0, 1):
                    if di == 0 and dj == 0 and dk == 0:
                        continue
                    nb = index.get((lev, li + di, lj + dj, lk + dk))
                    si, ti = _face_or_interior(di, ni)
                    if nb is not None:
                        out[mb, :nprim, tk, tj, ti] = uov[:, nb, sk, sj, si]
                        out[mb, nprim:, tk, tj, ti] = B[:, nb, sk, sj, si]
                    elif multilevel:
                        _fill_from_other_levels(out, uov, B, index, mb, lev, (li, lj, lk), (di, dj, dk))
    return out, index


def _fill_single_level_box(uov, B, loc, levels, index):
    """Fast path for a single-level mesh whose blocks tile a full box: stitch the blocks into one zero-padded
    global array per primitive and cut the ghost-padded blocks out of it.  Same result as the neighbour-by-
    neighbour copy (a ghost cell is the neighbouring block's edge cell, or zero outside the domain)."""
    if len(set(int(l) for l in levels)) != 1:
        return None
    nprim, nmb, nk, nj, ni = uov.shape
    lo = loc.min(axis=0)
    ext = loc.max(axis=0) - lo + 1
    if int(np.prod(ext)) != nmb or len(index) != nmb:
        return None
    n1, n2, n3 = (int(q) for q in ext)
    G = np.zeros((8, n3 * nk + 2, n2 * nj + 2, n1 * ni + 2))
    data = (uov, B)
    for mb in range(nmb):
        li, lj, lk = (int(q) for q in (loc[mb] - lo))
        sl = (slice(1 + lk * nk, 1 + (lk + 1) * nk), slice(1 + lj * nj, 1 + (lj + 1) * nj),
              slice(1 + li * ni, 1 + (li + 1) * ni))
        G[(slice(0, nprim),) + sl] = uov[:, mb]
        G[(slice(nprim, 8),) + sl] = B[:, mb]
    out = np.empty((nmb, 8, nk + 2, nj + 2, ni + 2))
    for mb in range(nmb):
        li, lj, lk = (int(q) for q in (loc[mb] - lo))
        out[mb] = G[:, lk * nk:(lk + 1) * nk + 2, lj * nj:(lj + 1) * nj + 2, li * ni:(li + 1) * ni + 2]
    return out


def _fill_from_other_levels(out, uov, B, index, mb, lev, loc, d):
    """Ghost cells across a refinement boundary (athenak.py:231-514).

    Coarser neighbour: injection (each ghost cell takes the value of the coarse cell that contains it).
    Finer neighbour: each ghost cell is the average of the 8 fine cells it covers, summed in the reference's order
    (so that faces and corners, where the reference is correct, agree with it bit for bit).
    """
    nprim = uov.shape[0]
    n = (uov.shape[4], uov.shape[3], uov.shape[2])          # (ni, nj, nk)
    data = np.concatenate([uov, B], axis=0)                  # (8, nmb, nk, nj, ni)
    # target cell index ranges (in padded coordinates) along each axis
    tgt = []
    for ax in range(3):
        if d[ax] == 1:
            tgt.append(np.array([n[ax] + 1]))
        elif d[ax] == -1:
            tgt.append(np.array([0]))
        else:
            tgt.append(np.arange(1, n[ax] + 1))
    # global fine-level cell coordinates of the target cells at this block's level
    gcell = [(loc[ax] * n[ax] + (tgt[ax] - 1)) for ax in range(3)]
    # --- coarser neighbour ---
    cl = tuple((loc[ax] + d[ax]) // 2 for ax in range(3))
    nb = index.get((lev - 1, cl[0], cl[1], cl[2]))
    if nb is not None:
        src = [np.floor_divide(gcell[ax], 2) - cl[ax] * n[ax] for ax in range(3)]
        if all(((s >= 0) & (s < n[ax])).all() for ax, s in enumerate(src)):
            out[mb][np.ix_(np.arange(8), tgt[2], tgt[1], tgt[0])] = data[:, nb][np.ix_(np.arange(8), src[2], src[1], src[0])]
        return
    # --- finer neighbours ---
    acc = np.zeros((8, len(tgt[2]), len(tgt[1]), len(tgt[0])))
    cnt = np.zeros((len(tgt[2]), len(tgt[1]), len(tgt[0])))
    # accumulation order of the reference (athenak.py:339-349, :405-422, :497-512): i offset outermost, k innermost
    for oi in (0, 1):
        for oj in (0, 1):
            for ok in (0, 1):
                off = (oi, oj, ok)
                fine = [2 * gcell[ax] + off[ax] for ax in range(3)]
                fb = [np.floor_divide(fine[ax], n[ax]) for ax in range(3)]
                fc = [fine[ax] - fb[ax] * n[ax] for ax in range(3)]
                for a, bk in enumerate(fb[2]):
                    for b_, bj in enumerate(fb[1]):
                        for c, bi in enumerate(fb[0]):
                            fm = index.get((lev + 1, int(bi), int(bj), int(bk)))
                            if fm is None:
                                continue
                            acc[:, a, b_, c] += data[:, fm, fc[2][a], fc[1][b_], fc[0][c]]
                            cnt[a, b_, c] += 1
    full = cnt == 8
    if full.any():
        vals = acc / 8.0
        kk, jj, ii = np.nonzero(full)
        out[mb][:, tgt[2][kk], tgt[1][jj], tgt[0][ii]] = vals[:, kk, jj, ii]


def build_block_grid(x1f, x2f, x3f, max_entries=1 << 26):
    """O(1) block lookup table for regular (power-of-two refined, axis-aligned) meshes.

    Returns ``(grid int32 (g3, g2, g1), gn, g0, ginv)`` or ``None`` when the mesh is irregular (the kernels
    then fall back to the reference's linear scan over meshblocks).
    """
    faces = [np.asarray(x1f), np.asarray(x2f), np.asarray(x3f)]
    lo = np.stack([f[:, 0] for f in faces])          # (3, nmb)
    hi = np.stack([f[:, -1] for f in faces])
    g0 = lo.min(axis=1)
    g1 = hi.max(axis=1)
    bw = (hi - lo).min(axis=1)
    gn = np.rint((g1 - g0) / bw).astype(np.int64)
    if np.any(gn < 1) or int(np.prod(gn)) > max_entries:
        return None
    clo = (lo - g0[:, None]) / bw[:, None]
    chi = (hi - g0[:, None]) / bw[:, None]
    if np.abs(clo - np.rint(clo)).max() > 1e-6 or np.abs(chi - np.rint(chi)).max() > 1e-6:
        return None
    clo = np.rint(clo).astype(np.int64)
    chi = np.rint(chi).astype(np.int64)
    grid = np.full((gn[2], gn[1], gn[0]), -1, dtype=np.int32)
    for mb in range(lo.shape[1]):
        grid[clo[2, mb]:chi[2, mb], clo[1, mb]:chi[1, mb], clo[0, mb]:chi[0, mb]] = mb
    return grid, gn.astype(np.int32), g0.astype(np.float64), (1.0 / bw).astype(np.float64)


class DeviceSampledFluidModel(GRMHDFluidModel):
    """Fluid models whose sampling runs in the CUDA kernels: subclasses provide ``snapshot()`` (an
    ``mk_snapshot`` handle) plus ``bhspin`` / ``fluid_gamma``."""

    _snap = None

    def snapshot(self, fill=True):
        raise NotImplementedError

    def release(self):
        if self._snap is not None:
            _cabi.call("mk_snapshot_destroy", self._snap)
            self._snap = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def get_prims_from_geodesics(self, S, profile=False):
        """athenak.py:527-637: interpolated primitives at S[..., :4]; zero outside the domain."""
        t0 = time.time()
        Sd = as_device(S)
        shape = tuple(Sd.shape[:-1])
        n = int(np.prod(shape)) if shape else 1
        out = empty((8, n))
        _cabi.call("mk_sample_prims", self.snapshot(), Sd.reshape(-1, 8), n, out, stream_ptr())
        names = ('dens', 'u', 'U1', 'U2', 'U3', 'B1', 'B2', 'B3')
        res = {k: DeviceArray.wrap(out[i].reshape(shape)) for i, k in enumerate(names)}
        if profile:
            __import__('torch').cuda.synchronize()
            print(f"Time to compute meshblock indices and primitives (one kernel): {time.time() - t0}")
        return res

    def get_fluid_scalars_from_geodesics(self, S, fallback_pitch_angle=np.pi / 3., profile=False):
        """athenak.py:639-812: dens, u, pitch_angle, kdotu, b at the points of S (nsteps, npx, 8)."""
        t0 = time.time()
        Sd = as_device(S)
        shape = tuple(Sd.shape[:-1])
        n = int(np.prod(shape)) if shape else 1
        out = empty((5, n))
        _cabi.call("mk_sample_scalars", self.snapshot(), float(self.bhspin), Sd.reshape(-1, 8), n,
                   float(fallback_pitch_angle), out, stream_ptr())
        names = ('dens', 'u', 'pitch_angle', 'kdotu', 'b')
        res = {k: DeviceArray.wrap(out[i].reshape(shape)) for i, k in enumerate(names)}
        if profile:
            __import__('torch').cuda.synchronize()
            print(f"Time to compute meshblock indices and scalar data (one kernel): {time.time() - t0}")
        return res


class AnalyticTorusFluidModel(DeviceSampledFluidModel):
    """BASELINE cfg3: analytic Keplerian thin torus (power-law density, toroidal field at fixed plasma beta),
    evaluated in closed form at every sample point instead of interpolating snapshot cells.  The reference
    ships no analytic model; this is a ``GRMHDFluidModel`` duck type pushed through the same fluid-frame,
    thermodynamics and transfer code.  Zero outside the sphere ``r <= r_out``."""

    def __init__(self, bhspin, fluid_gamma=13. / 9, R0=8.0, R_in=2.5, p=1.5, h=0.3, u0=0.25, beta0=3.0,
                 dens_scale=1.0, r_out=40.0):
        self.bhspin, self.fluid_gamma = bhspin, fluid_gamma
        self.params = (fluid_gamma, R0, R_in, p, h, u0, beta0, dens_scale, r_out)
        self._snap = None
        self.storage, self.lookup = "analytic", "analytic"

    def snapshot(self, fill=True):
        require_gpu()
        if self._snap is None:
            handle = ctypes.c_void_p()
            _cabi.call("mk_snapshot_create_torus", (ctypes.c_double * 9)(*[float(q) for q in self.params]),
                       ctypes.byref(handle))
            self._snap = handle
        return self._snap

    def snapshot_bytes(self):
        return 0


class AthenakFluidModel(DeviceSampledFluidModel):

    def __init__(self, grmhd_filename, bhspin, fluid_gamma=None):
        """athenak.py:50-53.  Reads an AthenaK ``.athdf`` dump (with h5py if installed, else with the built-in
        minimal HDF5 reader) or an ``.npz`` file holding the same datasets (``scripts/athdf_to_npz.py``)."""
        if str(grmhd_filename).endswith(".npz"):
            arrays = self._read_npz(grmhd_filename)
        else:
            arrays = self._read_athdf(grmhd_filename)
        self._setup(bhspin=bhspin, fluid_gamma=fluid_gamma, **arrays)

    @classmethod
    def from_arrays(cls, uov, B, x1v, x2v, x3v, x1f, x2f, x3f, LogicalLocations, Levels, bhspin,
                    fluid_gamma=None, VariableNames=('dens', 'velx', 'vely', 'velz', 'eint', 'bcc1', 'bcc2', 'bcc3'),
                    storage='auto', lookup='auto', ghost_fill='device'):
        """Construct from the arrays of an ``.athdf`` file (athenak.py:79-103) without touching disk.

        storage: 'f64', 'f32' or 'auto' (float32 cells when every value is float32-representable —
        AthenaK writes float32 — which halves the sampling traffic without changing a single bit).
        lookup: 'grid' (O(1) block table), 'scan' (the reference's linear scan) or 'auto'.
        ghost_fill: 'device' (ghost zones filled by the CUDA kernel that repacks the snapshot; only the interior
        arrays are uploaded) or 'host' (NumPy fill + upload of the padded blocks, the reference's route).
        """
        self = cls.__new__(cls)
        self._setup(uov=uov, B=B, x1v=x1v, x2v=x2v, x3v=x3v, x1f=x1f, x2f=x2f, x3f=x3f,
                    LogicalLocations=LogicalLocations, Levels=Levels, VariableNames=VariableNames,
                    bhspin=bhspin, fluid_gamma=fluid_gamma, storage=storage, lookup=lookup, ghost_fill=ghost_fill)
        return self

    @classmethod
    def replica(cls, x1v, x2v, x3v, x1f, x2f, x3f, bhspin, fluid_gamma, VariableNames, block_shape, storage):
        """Geometry-only model for a rank that receives the snapshot cells from another rank
        (``multigpu.replicate_snapshot``): no host copy of the primitives is kept."""
        self = cls.__new__(cls)
        self.bhspin, self.fluid_gamma = bhspin, fluid_gamma
        self.variable_names = np.array(list(VariableNames))
        self._uov = self._B = self._amb = None
        self._ghost_fill = 'host'
        self._block_shape = tuple(int(q) for q in block_shape)
        self.x1v, self.x2v, self.x3v = (np.asarray(q, dtype=np.float64) for q in (x1v, x2v, x3v))
        self.x1f, self.x2f, self.x3f = (np.asarray(q, dtype=np.float64) for q in (x1f, x2f, x3f))
        self.nprim_all = 8
        self._storage, self._lookup = storage, 'auto'
        self._snap, self._snap_device = None, None
        return self

    def replica_meta(self):
        """What another rank needs to build a ``replica`` of this model."""
        return dict(x1v=self.x1v, x2v=self.x2v, x3v=self.x3v, x1f=self.x1f, x2f=self.x2f, x3f=self.x3f,
                    bhspin=self.bhspin, fluid_gamma=self.fluid_gamma, VariableNames=list(self.variable_names),
                    block_shape=self._block_shape,
                    storage=self.storage)

    DATASETS = ('x1v', 'x2v', 'x3v', 'x1f', 'x2f', 'x3f', 'uov', 'B', 'LogicalLocations', 'Levels')

    @classmethod
    def _read_npz(cls, filename):
        with np.load(filename, allow_pickle=False) as z:
            out = {k: np.array(z[k]) for k in cls.DATASETS}
            out['VariableNames'] = [str(n) for n in z['VariableNames']]
        return out

    @staticmethod
    def _read_athdf(filename):
        """athenak.py:79-103.  Uses h5py when it is installed, else the built-in minimal HDF5 reader
        (``_hdf5_min``: the subset of the format h5py writes for AthenaK dumps)."""
        try:
            import h5py
        except ImportError:
            from ._hdf5_min import read_athdf
            return read_athdf(filename)
        with h5py.File(filename, 'r') as hfp:      # pragma: no cover - h5py is not installed in the build image
            out = {k: np.array(hfp[k]) for k in ('x1v', 'x2v', 'x3v', 'x1f', 'x2f', 'x3f', 'uov', 'B',
                                                 'LogicalLocations', 'Levels')}
            out['VariableNames'] = [n.decode('utf-8') if isinstance(n, bytes) else str(n)
                                    for n in hfp.attrs['VariableNames']]
        return out

    def _setup(self, uov, B, x1v, x2v, x3v, x1f, x2f, x3f, LogicalLocations, Levels, VariableNames, bhspin,
               fluid_gamma, storage='auto', lookup='auto', ghost_fill='device'):
        if ghost_fill not in ('device', 'host'):
            raise ValueError("ghost_fill must be 'device' or 'host'")
        self.bhspin = bhspin
        self.fluid_gamma = fluid_gamma
        self.variable_names = np.array(list(VariableNames))
        # interior arrays as the file holds them; float32 input (what AthenaK writes) is kept as float32
        keep = lambda a: np.asarray(a) if np.asarray(a).dtype == np.float32 else np.asarray(a, dtype=np.float64)
        self._uov, self._B = keep(uov), keep(B)
        if self._uov.ndim != 5 or self._B.ndim != 5 or self._uov.shape[1:] != self._B.shape[1:]:
            raise ValueError("uov and B must have shapes (nvar, nmb, nk, nj, ni) with equal block shapes")
        if self._uov.shape[0] + self._B.shape[0] != 8:
            raise ValueError("uov and B must hold 8 primitives together")
        nmb, nk, nj, ni = self._uov.shape[1:]
        self._block_shape = (nmb, 8, nk + 2, nj + 2, ni + 2)
        self._amb = None
        self._ghost_fill = ghost_fill
        self.x1v, self.x2v, self.x3v = (np.asarray(q, dtype=np.float64) for q in (x1v, x2v, x3v))
        self.x1f, self.x2f, self.x3f = (np.asarray(q, dtype=np.float64) for q in (x1f, x2f, x3f))
        self.Levels = np.asarray(Levels)
        self.LogicalLocations = np.asarray(LogicalLocations)
        self.mb_index_map = {(int(self.Levels[mb]),) + tuple(int(q) for q in self.LogicalLocations[mb]): mb
                             for mb in range(nmb)}
        self.nprim_all = 8
        self._storage = storage
        self._lookup = lookup
        self._snap = None
        self._snap_device = None

    @property
    def all_meshblocks(self):
        """The reference's ghost-padded ``(nmb, 8, nk+2, nj+2, ni+2)`` array (athenak.py:105-158), built on the
        host on first access.  The device snapshot does not need it (``ghost_fill='device'``)."""
        if self._amb is None and self._uov is not None:
            self._amb, _ = fill_ghost_zones(self._uov, self._B, self.LogicalLocations, self.Levels)
        return self._amb

    def device_meshblocks(self):
        """``all_meshblocks`` as reconstructed from the device snapshot cells (``mk_snapshot_unpack``)."""
        out = empty(self._block_shape)
        _cabi.call("mk_snapshot_unpack", self.snapshot(), (ctypes.c_int * 8)(*self._prim_index()), out, stream_ptr())
        return DeviceArray.wrap(out)

    def get_index_for_primitive_by_name(self, prim):
        """athenak.py:55-65."""
        prim = prim.lower().strip()
        names = [v.lower().strip() for v in self.variable_names]
        return names.index(prim) if prim in names else -1

    # ---- device snapshot ---------------------------------------------------------------------
    def _prim_index(self):
        idx = [self.get_index_for_primitive_by_name(p) for p in CANONICAL_PRIMS]
        iU1, iU2, iU3 = idx[2:5]
        iB1, iB2, iB3 = idx[5:8]
        if iU2 != iU1 + 1 or iU3 != iU1 + 2:
            raise ValueError("Velocity indices are not as expected")          # athenak.py:707-708
        if iB2 != iB1 + 1 or iB3 != iB1 + 2:
            raise ValueError("Magnetic field indices are not as expected")    # athenak.py:709-710
        if min(idx) < 0:
            raise ValueError(f"snapshot lacks one of the primitives {CANONICAL_PRIMS}")
        return idx

    def snapshot(self, fill=True):
        """The device-resident repacked snapshot handle (created on first use, reused afterwards).
        ``fill=False`` allocates the cells without uploading them (a replica filled by a broadcast)."""
        dev = require_gpu()
        if self._snap is not None and self._snap_device == dev:
            return self._snap
        import time
        torch = __import__('torch')
        t_start = time.perf_counter()
        nmb, _, nk2, nj2, ni2 = self._block_shape
        if self._uov is None and fill:
            raise ValueError("a replica model has no host data: fill it with multigpu.replicate_snapshot")
        on_device = fill and self._ghost_fill == 'device'
        storage = self._storage
        amb = None
        if fill and not on_device:
            amb = self.all_meshblocks
            if storage == 'auto':
                storage = 'f32' if np.array_equal(amb.astype(np.float32).astype(np.float64), amb) else 'f64'
        geom = np.stack([self.x1f[:, 0], self.x2f[:, 0], self.x3f[:, 0],
                         self.x1f[:, -1], self.x2f[:, -1], self.x3f[:, -1],
                         self.x1v[:, 0], self.x2v[:, 0], self.x3v[:, 0],
                         self.x1v[:, 1] - self.x1v[:, 0], self.x2v[:, 1] - self.x2v[:, 0],
                         self.x3v[:, 1] - self.x3v[:, 0]]).T.copy()        # (nmb, 12) records
        bbox_lo = (ctypes.c_double * 3)(self.x1f[:, 0].min(), self.x2f[:, 0].min(), self.x3f[:, 0].min())
        bbox_hi = (ctypes.c_double * 3)(self.x1f[:, -1].max(), self.x2f[:, -1].max(), self.x3f[:, -1].max())
        grid = None
        if self._lookup in ('auto', 'grid'):
            grid = build_block_grid(self.x1f, self.x2f, self.x3f)
            if grid is None and self._lookup == 'grid':
                raise ValueError("mesh is not regular enough for the block-grid lookup; use lookup='scan'")
        pidx = (ctypes.c_int * 8)(*self._prim_index())
        handle = ctypes.c_void_p()
        d_geom = as_device(geom)
        if grid is not None:
            g, gn, g0, ginv = grid
            d_grid = as_device(g, dtype=__import__('torch').int32)
            c_gn = (ctypes.c_int * 3)(*[int(q) for q in gn])
            c_g0 = (ctypes.c_double * 3)(*[float(q) for q in g0])
            c_gi = (ctypes.c_double * 3)(*[float(q) for q in ginv])
        else:
            d_grid, c_gn, c_g0, c_gi = None, None, None, None
        t_prep = t_up = time.perf_counter()
        if on_device:
            src_f32 = self._uov.dtype == np.float32 and self._B.dtype == np.float32
            dt = torch.float32 if src_f32 else torch.float64
            d_uov, d_B = self._upload_meshblocks(self._uov, dtype=dt), self._upload_meshblocks(self._B, dtype=dt)
            loc = np.ascontiguousarray(self.LogicalLocations, dtype=np.int32)
            lev = np.ascontiguousarray(self.Levels, dtype=np.int32)
            stored = ctypes.c_int(0)
            torch.cuda.current_stream().synchronize()
            t_up = time.perf_counter()
            _cabi.call("mk_snapshot_create_from_interiors", nmb, nk2 - 2, nj2 - 2, ni2 - 2, d_uov,
                       int(self._uov.shape[0]), d_B, int(self._B.shape[0]), 1 if src_f32 else 0, pidx,
                       loc.ctypes.data, lev.ctypes.data, d_geom, d_grid, c_gn, c_g0, c_gi, bbox_lo, bbox_hi,
                       {'f64': 0, 'f32': 1, 'auto': 2}[storage], ctypes.byref(stored), ctypes.byref(handle),
                       stream_ptr())
            storage = 'f32' if stored.value else 'f64'
            del d_uov, d_B
        else:
            d_mb = self._upload_meshblocks(amb) if fill else None
            if storage == 'auto':          # a replica: the caller passes the storage of the source rank
                storage = 'f64'
            _cabi.call("mk_snapshot_create", nmb, nk2 - 2, nj2 - 2, ni2 - 2, d_mb, pidx, d_geom, d_grid, c_gn, c_g0,
                       c_gi, bbox_lo, bbox_hi, 1 if storage == 'f32' else 0, ctypes.byref(handle), stream_ptr())
            del d_mb
        torch.cuda.current_stream().synchronize()
        t_end = time.perf_counter()
        # phases of the set-up in ms: host-side geometry / lookup tables, upload of the cell arrays, kernel
        # (ghost fill + repack; includes the padded-array upload when ghost_fill='host')
        self.setup_timing = dict(host_prep=1e3 * (t_prep - t_start), upload=1e3 * (t_up - t_prep),
                                 ghost_fill=1e3 * (t_end - t_up))
        self._snap = handle
        self._snap_device = dev
        self.storage = storage
        self.lookup = 'grid' if grid is not None else 'scan'
        return self._snap

    @staticmethod
    def _upload_meshblocks(amb, chunk_bytes=256 << 20, dtype=None):
        """Host -> device copy of a large block array from pageable memory (``mk_upload_pageable``): pinning a
        multi-GB NumPy array first costs more than the copy itself, and the driver's own staging of a pageable
        cudaMemcpy runs on one thread."""
        torch = __import__('torch')
        dtype = dtype or torch.float64
        if amb.nbytes <= (16 << 20):
            return as_device(amb, dtype=dtype)
        np_dtype = {torch.float64: np.float64, torch.float32: np.float32}[dtype]
        h = np.ascontiguousarray(amb, dtype=np_dtype).reshape(-1)
        d = empty(amb.shape, dtype=dtype)
        # native pipelined uploader: host threads fill one pinned staging buffer while the DMA engine empties the other
        _cabi.call("mk_upload_pageable", d, h.ctypes.data, int(h.nbytes), 0, stream_ptr())
        return d

    def snapshot_bytes(self):
        return int(_cabi.load().mk_snapshot_bytes(self.snapshot()))
