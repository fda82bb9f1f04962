"""Golden vectors produced by the reference's own athenak.py and images.py (tests/golden/reference_fluid_golden.npz).

Third member of the family (make_reference_golden.py, make_reference_geodesics_golden.py): here the WHOLE reference
package is imported from /root/reference -- its real __init__, geodesics.py, transfer.py, electrons.py, images.py,
grmhd/grmhd.py, grmhd/athenak.py, all unmodified -- with two stand-ins in sys.modules:

    jax      NumPy-backed (see make_reference_geodesics_golden.py; jacfwd = complex step), plus jax.config / jax.lib
    h5py     File(name) serves in-memory arrays registered under that name (the loader's only I/O)

so that the reference's own loader (ghost-zone fill incl. refinement boundaries), its own
get_prims_from_geodesics / get_fluid_scalars_from_geodesics and its own make_image run on synthetic AthenaK-shaped
snapshots.  Frozen: all_meshblocks of a single-level and of a two-level mesh, sampled primitives and fluid scalars
along reference trajectories, and two 6x6 images (one chunked with max_chunk_bytes).  Run in the build container
(about five minutes), commit the .npz.

    python tests/golden/make_reference_fluid_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import make_reference_geodesics_golden as G  # noqa: E402
from make_reference_golden import _Arr, arr  # noqa: E402

OUT = os.path.join(HERE, "reference_fluid_golden.npz")
FILES = {}


class _FakeH5File:
    def __init__(self, name, mode='r'):
        self.d = FILES[name]
        self.attrs = {"VariableNames": [n.encode() for n in self.d["VariableNames"]]}

    def __getitem__(self, k):
        return self.d[k]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass


def _scan_any(f, init, xs):
    """lax.scan with arbitrary carry (None here) and a single stacked output"""
    carry, ys = init, []
    for x in xs:
        carry, y = f(carry, x)
        ys.append(None if y is None else np.asarray(y))
    return carry, (None if (not ys or ys[0] is None) else np.stack(ys).view(_Arr))


def _jax_getitem(self, key):
    """jax.Array indexing with integer arrays: negative indices wrap, out-of-range indices are CLAMPED (NumPy raises).
    athenak.py:594-601 relies on it for points outside the domain, whose garbage values it zeroes afterwards."""
    if isinstance(key, tuple) and any(isinstance(k, np.ndarray) and k.dtype.kind in "iu" for k in key):
        new = []
        for ax, k in enumerate(key):
            if isinstance(k, np.ndarray) and k.dtype.kind in "iu":
                n = self.shape[ax]
                k = np.asarray(k)
                k = np.where(k < 0, k + n, k)
                k = np.clip(k, 0, n - 1)
            new.append(k)
        key = tuple(new)
    return np.ndarray.__getitem__(self, key)


def install():
    _Arr.__getitem__ = _jax_getitem
    G.install_stand_in()
    sys.modules.pop("mahakala", None)                       # import the REAL package this time
    jax = sys.modules["jax"]
    jnp, lax = jax.numpy, jax.lax
    for name in ("einsum", "stack", "ones_like", "cos", "arccos", "ones", "sin", "exp"):
        setattr(jnp, name, G._wrap(getattr(np, name)))
    geo_scan = lax.scan

    def scan(f, init, xs):                                  # tuple outputs with an array carry: the geodesic scan
        return geo_scan(f, init, xs) if init is not None and not isinstance(init, type(None)) and \
            getattr(init, "ndim", 0) == 2 else _scan_any(f, init, xs)
    lax.scan = scan
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    lib = types.ModuleType("jax.lib")
    lib.xla_bridge = types.SimpleNamespace(get_backend=lambda: types.SimpleNamespace(platform="NumPy stand-in"))
    jax.lib = lib
    sys.modules["jax.lib"] = lib
    h5 = types.ModuleType("h5py")
    h5.File = _FakeH5File
    sys.modules["h5py"] = h5
    sys.path.insert(0, "/root/reference")


def register(name, a):
    FILES[name] = dict(a)


def compute():
    install()
    import mahakala as ma                                    # the reference itself
    from mahakala.grmhd.athenak import AthenakFluidModel
    from mahakala.images import make_image
    from helpers import two_level_mesh
    from mahakala_b200.synthetic import make_synthetic_snapshot
    out = {}
    a = 0.94
    # ---- loader: single level (2x2x2 blocks of 8^3, dx = 2) and two levels (root block (1,1,1) refined) ----
    single = make_synthetic_snapshot(ncells=16, block=8, extent=16.0, seed=0)
    register("single.athdf", single)
    model = AthenakFluidModel("single.athdf", a, fluid_gamma=single["fluid_gamma"])
    out["single_all_meshblocks"] = np.asarray(model.all_meshblocks)
    amr, _ = two_level_mesh(n=8)
    register("amr.athdf", amr)
    amr_model = AthenakFluidModel("amr.athdf", a, fluid_gamma=amr["fluid_gamma"])
    out["amr_all_meshblocks"] = np.asarray(amr_model.all_meshblocks)
    # ---- sampling along reference trajectories (the 6x6 bundle of reference_geodesics_golden.npz) ----
    geo = np.load(os.path.join(HERE, "reference_geodesics_golden.npz"))
    S = geo["traj_S"][200:520:4]                              # (80, 36, 8): far outside -> through the box -> horizon
    out["sample_S"] = S
    prims = model.get_prims_from_geodesics(arr(S))
    out["sample_prims"] = np.stack([np.asarray(prims[k]) for k in ('dens', 'u', 'U1', 'U2', 'U3', 'B1', 'B2', 'B3')])
    sc = model.get_fluid_scalars_from_geodesics(arr(S))
    out["sample_scalars"] = np.stack([np.asarray(sc[k]) for k in ('dens', 'u', 'pitch_angle', 'kdotu', 'b')])
    pts = arr(np.concatenate([S[:, :5].reshape(-1, 8)[:, :4]]))
    # ---- the reference's make_image on the single-level snapshot: one pass and chunked (12 pixels per chunk) ----
    out["image_res6"] = np.asarray(make_image(model, resolution=6))
    out["image_res6_chunked"] = np.asarray(make_image(model, resolution=6, max_chunk_bytes=12 * 4 * 20 * 10000))
    out["image_res6_345GHz_i30"] = np.asarray(make_image(model, camera_inclination=30, observing_frequency=345e9,
                                                         r_high=10, resolution=6, max_nsteps=3000))
    return out


if __name__ == "__main__":
    res = compute()
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, {k: np.asarray(v).shape for k, v in res.items()})
