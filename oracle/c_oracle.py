"""ctypes front-end of the C oracle (oracle/mk_oracle.c).  TEST INFRASTRUCTURE ONLY — see the header of
mk_oracle.c.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    so = os.path.join(_HERE, "libmk_oracle.so")
    src = os.path.join(_HERE, "mk_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def num_threads():
    return int(lib().orc_num_threads())


def use_all_cores():
    """Run the OpenMP loops on every core this process may use (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


def integrate(N, s0, div, tol, bhspin, dump=False):
    """-> dict(final (npx,8), nsteps (npx,), r_last (npx,)[, S (nrows,npx,8), dt (nrows,npx)]).

    nrows follows geodesics.py:275-281: first all-zero row + 2 (N if there is none / it is row 0)."""
    s0 = np.ascontiguousarray(s0, dtype=np.float64)
    npx = s0.shape[0]
    final = np.empty((npx, 8))
    nsteps = np.empty(npx, dtype=np.int32)
    r_last = np.empty(npx)
    L = lib()
    L.orc_integrate(ctypes.c_int(N), ctypes.c_int(npx), _d(s0), ctypes.c_double(div), ctypes.c_double(tol),
                    ctypes.c_double(bhspin), _d(final), nsteps.ctypes.data_as(_ip), _d(r_last),
                    None, None, ctypes.c_int(0))
    out = dict(final=final, nsteps=nsteps, r_last=r_last)
    if dump:
        M = int(nsteps.max()) if npx else 0
        first_zero = M if (1 <= M <= N - 1) else N
        nrows = min(first_zero + 2, N)
        S = np.empty((nrows, npx, 8))
        dt = np.empty((nrows, npx))
        L.orc_integrate(ctypes.c_int(N), ctypes.c_int(npx), _d(s0), ctypes.c_double(div),
                        ctypes.c_double(tol), ctypes.c_double(bhspin), None, None, None,
                        _d(S), _d(dt), ctypes.c_int(nrows))
        out.update(S=S, dt=dt)
    return out


def geodesic_integrator(N, s0, div, tol, bhspin):
    """Drop-in for mahakala_oracle.geodesic_integrator (same outputs), ~1000x faster."""
    o = integrate(N, s0, div, tol, bhspin, dump=True)
    return o["S"], o["dt"]


def rhs(state, bhspin):
    state = np.ascontiguousarray(state, dtype=np.float64)
    out = np.empty_like(state)
    lib().orc_rhs(ctypes.c_int(state.shape[0]), _d(state), ctypes.c_double(bhspin), _d(out))
    return out


def _torus_args(model):
    tp = np.ascontiguousarray(model.params_array(), dtype=np.float64)
    z = np.zeros(4)
    args = [ctypes.c_int(0)] * 4 + [_d(z)] * 7
    return args, [tp, z], tp


def _snap_args(model):
    """model: oracle AthenakFluidModel (reference layout arrays)."""
    if hasattr(model, "params_array"):
        args, keep, tp = _torus_args(model)
        return args, keep
    d = np.ascontiguousarray(model.all_meshblocks, dtype=np.float64)
    nmb, _, nk2, nj2, ni2 = d.shape
    arrs = [np.ascontiguousarray(q, dtype=np.float64) for q in
            (model.x1f, model.x2f, model.x3f, model.x1v, model.x2v, model.x3v)]
    args = [ctypes.c_int(nmb), ctypes.c_int(nk2 - 2), ctypes.c_int(nj2 - 2), ctypes.c_int(ni2 - 2), _d(d)]
    args += [_d(q) for q in arrs]
    return args, [d] + arrs


def sample(model, S, mode="scalars", fallback_pitch_angle=np.pi / 3.):
    """S (..., 8) -> dict of arrays shaped S.shape[:-1] (get_fluid_scalars / get_prims semantics)."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    shape = S.shape[:-1]
    n = int(np.prod(shape))
    nout = 5 if mode == "scalars" else 8
    out = np.empty((nout, n))
    args, keep = _snap_args(model)
    torus = _d(keep[0]) if hasattr(model, "params_array") else None
    lib().orc_sample(ctypes.c_int(0 if mode == "scalars" else 1), ctypes.c_long(n), _d(S), *args,
                     ctypes.c_double(model.bhspin), ctypes.c_double(fallback_pitch_angle), torus, _d(out))
    names = (("dens", "u", "pitch_angle", "kdotu", "b") if mode == "scalars"
             else ("dens", "u", "U1", "U2", "U3", "B1", "B2", "B3"))
    return {k: out[i].reshape(shape) for i, k in enumerate(names)}


def set_sigma_cut(cut=100.):
    """images.py:116 uses 100; other values only for what-if checks of fixtures (does the cut matter here?)."""
    lib().orc_set_sigma_cut(ctypes.c_double(cut))


def emission(S, prims, bhspin, fluid_gamma, r_high, units, nu_obs):
    """athenak.py:760-794 + images.py:87-118 on arbitrary (state, primitives) pairs: S (n, 8), prims (n, 8) in file
    order dens, velx, vely, velz, eint, bcc1..3 -> invariant em, ab (nfreq, n) and sigma (n,)."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    prims = np.ascontiguousarray(prims, dtype=np.float64)
    nu = np.ascontiguousarray(np.atleast_1d(nu_obs), dtype=np.float64)
    n = S.shape[0]
    em = np.empty((nu.size, n))
    ab = np.empty((nu.size, n))
    sigma = np.empty(n)
    lib().orc_emission(ctypes.c_long(n), _d(S), _d(prims), ctypes.c_double(bhspin), ctypes.c_double(fluid_gamma),
                       ctypes.c_double(r_high), ctypes.c_double(units["Ne_unit"]), ctypes.c_double(units["B_unit"]),
                       ctypes.c_int(nu.size), _d(nu), _d(em), _d(ab), _d(sigma))
    return em, ab, sigma


def render(model, s0, units, nu_obs, r_high=40., N=10000, div=40., tol=1e-4):
    """images.py:56-144 per ray -> (image (nfreq, npx), nsteps (npx,), n_in_domain)."""
    s0 = np.ascontiguousarray(s0, dtype=np.float64)
    npx = s0.shape[0]
    nu = np.ascontiguousarray(np.atleast_1d(nu_obs), dtype=np.float64)
    img = np.empty((nu.size, npx))
    nsteps = np.empty(npx, dtype=np.int32)
    nin = ctypes.c_int64(0)
    args, keep = _snap_args(model)
    lib().orc_render(ctypes.c_int(N), ctypes.c_int(npx), _d(s0), ctypes.c_double(div), ctypes.c_double(tol),
                     *args, ctypes.c_double(model.bhspin), ctypes.c_double(model.fluid_gamma),
                     ctypes.c_double(r_high), ctypes.c_double(units["Ne_unit"]),
                     ctypes.c_double(units["B_unit"]), ctypes.c_double(units["L_unit"]),
                     ctypes.c_int(nu.size), _d(nu), _d(keep[0]) if hasattr(model, "params_array") else None,
                     _d(img), nsteps.ctypes.data_as(_ip), ctypes.byref(nin))
    return img, nsteps, int(nin.value)
