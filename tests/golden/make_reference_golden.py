"""Golden vectors produced by the REFERENCE'S OWN SOURCE FILES (tests/golden/reference_golden.npz).

JAX is not installable in this image, so the reference package cannot be imported as a whole.  Four of its modules,
however, use JAX only as an array library:

    /root/reference/mahakala/constants.py        cgs constants
    /root/reference/mahakala/electrons.py        rlow_rhigh_model            (plain arithmetic)
    /root/reference/mahakala/grmhd/grmhd.py      GRMHDFluidModel.get_units   (plain arithmetic + numpy)
    /root/reference/mahakala/transfer.py         synchrotron_coefficients, solve_specific_intensity,
                                                 solve_attenuated_emissivity (jnp elementwise, .at[].set, lax.select,
                                                 lax.scan)

This script executes those files UNMODIFIED, straight from /root/reference, against a ~40-line NumPy stand-in for the
`jax.numpy` / `jax.lax` names they touch (float64 throughout, as the reference runs with jax_enable_x64), and freezes
inputs and outputs.  The vectors therefore state what the reference's text computes under IEEE double arithmetic with
NumPy's libm; XLA's own exp / pow / sin may differ from them in the last bits, which is why consumers compare at
1e-12, not bit for bit.  /root/reference exists only in the build container: run this here, commit the .npz.

    python tests/golden/make_reference_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/mahakala"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")


# ---- the stand-in: just enough of jax.numpy / jax.lax for transfer.py -------------------------------------------
class _Arr(np.ndarray):
    """ndarray with jax's functional update syntax x.at[idx].set(v)"""

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def set(self, v):
        out = np.array(self.a, copy=True).view(_Arr)
        out[self.idx] = v
        return out


def arr(x):
    return np.asarray(x, dtype=np.float64).view(_Arr)


def _scan(f, init, xs):
    carry, ys = init, []
    for x in xs:
        carry, y = f(np.array(carry, copy=True).view(_Arr), x)     # the reference's body updates its carry in place
        ys.append(y)
    return carry, (None if ys and ys[0] is None else np.stack(ys).view(_Arr))


def install_stand_in():
    jnp = types.ModuleType("jax.numpy")
    for name in ("sin", "exp", "sqrt", "isnan", "arange", "pi"):
        setattr(jnp, name, getattr(np, name))
    jnp.zeros = lambda n: np.zeros(n).view(_Arr)
    lax = types.ModuleType("jax.lax")
    lax.select = lambda pred, a, b: np.where(pred, a, b).view(_Arr)
    lax.scan = _scan
    jax = types.ModuleType("jax")
    jax.numpy, jax.lax = jnp, lax
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.lax": lax})
    pkg = types.ModuleType("mahakala")          # bare package object: the real __init__ imports the whole of JAX
    pkg.__path__ = [REF]
    sys.modules["mahakala"] = pkg


def load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def compute():
    install_stand_in()
    constants = load("mahakala.constants", "constants.py")
    electrons = load("mahakala.electrons", "electrons.py")
    grmhd = load("mahakala.grmhd.grmhd", "grmhd/grmhd.py")
    transfer = load("mahakala.transfer", "transfer.py")
    out = {}
    for k in ("EE", "KB", "CL", "ME", "HPL", "GNEWT", "Msun"):
        out["const_" + k] = np.float64(getattr(constants, k))
    for k in ("MP", "EC"):
        out["const_" + k] = np.float64(getattr(electrons, k))
    # get_units (grmhd.py:40-55) for the demo notebooks' parameters and a second set
    for tag, (M, ms) in {"a": (6.2e9 * constants.Msun, 1.e26), "b": (4.1e6 * constants.Msun, 3.7e17)}.items():
        u = grmhd.GRMHDFluidModel().get_units(M, ms)
        out["units_%s_in" % tag] = np.array([M, ms])
        out["units_%s_out" % tag] = np.array([u[k] for k in ("L_unit", "T_unit", "dens_unit", "Ne_unit", "B_unit")])
    rng = np.random.default_rng(2024)
    n = 4096
    # rlow_rhigh_model (electrons.py:32-50), incl. zeros / NaN inputs as images.py:87-95 feeds them
    dens = np.exp(rng.normal(-2, 2, n)); u = np.exp(rng.normal(-4, 2, n)); beta = np.exp(rng.normal(0, 3, n))
    dens[:8] = 0.0; u[4:12] = 0.0; beta[10:16] = np.nan; beta[16:20] = np.inf
    with np.errstate(all="ignore"):
        out["theta_in"] = np.stack([dens, u, beta])
        out["theta_default"] = np.asarray(electrons.rlow_rhigh_model(arr(dens), arr(u), arr(beta)))
        out["theta_r10_r160"] = np.asarray(electrons.rlow_rhigh_model(arr(dens), arr(u), arr(beta), r_low=10, r_high=160))
        # synchrotron_coefficients (transfer.py:30-86): broad log-normal inputs + the special cases of the NaN clean-up
        Ne = np.exp(rng.normal(12, 3, n)); Th = np.exp(rng.normal(1, 2, n)); B = np.exp(rng.normal(1, 2, n))
        pitch = rng.uniform(0, np.pi, n); nu = 230e9 * np.exp(rng.normal(0, 1, n))
        Th[:6] = [0.0, 0.29999, 0.3, np.nan, np.inf, 1e-300]; B[6:10] = [0.0, np.nan, 1e-30, 1e30]
        pitch[10:14] = [0.0, np.pi, np.nan, np.pi / 3.]; Ne[14:17] = [0.0, np.nan, np.inf]; nu[17:20] = [0.0, 1e3, 1e25]
        Th[20:40] = np.exp(rng.uniform(np.log(5e2), np.log(5e4), 20))       # bx < 2e-3: the series branch of B_nu
        out["syn_in"] = np.stack([Ne, Th, B, pitch, nu])
        for tag, kw in {"inv": dict(invariant=True, rescale_nu=1. / 230e9), "inv1": dict(invariant=True),
                        "plain": dict(invariant=False)}.items():
            em, ab = transfer.synchrotron_coefficients(arr(Ne), arr(Th), arr(B), arr(pitch), arr(nu), **kw)
            out["syn_em_" + tag], out["syn_ab_" + tag] = np.asarray(em), np.asarray(ab)
        # transfer solvers (transfer.py:89-144) on a ragged bundle: dt < 0 then 0 once a ray has frozen
        nrows, npx = 257, 96
        # alpha |dt| L ~ 0.03 (max ~ 2): the regime in which the explicit-Euler update is stable and finite
        em2 = np.exp(rng.normal(-40, 3, (nrows, npx))); ab2 = np.exp(rng.normal(-38, 1, (nrows, npx)))
        em2[rng.random((nrows, npx)) < 0.3] = 0.0; ab2[em2 == 0.0] = 0.0
        dt = -np.exp(rng.normal(0, 1, (nrows, npx)))
        stop = rng.integers(1, nrows, npx)
        dt[np.arange(nrows)[:, None] >= stop[None, :]] = 0.0
        L = 9.157e14
        out["tr_in_em"], out["tr_in_ab"], out["tr_in_dt"], out["tr_in_L"] = em2, ab2, dt, np.float64(L)
        out["tr_I"] = np.asarray(transfer.solve_specific_intensity(arr(em2), arr(ab2), arr(dt), L))
        I2, dIs = transfer.solve_specific_intensity(arr(em2), arr(ab2), arr(dt), L, dIs=True)
        out["tr_I_dIs"], out["tr_dIs"] = np.asarray(I2), np.asarray(dIs)
        out["tr_attenuated"] = np.asarray(transfer.solve_attenuated_emissivity(arr(em2), arr(ab2), arr(dt), L))
    return out


if __name__ == "__main__":
    res = compute()
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, {k: np.asarray(v).shape for k, v in res.items()})
