"""Golden vectors produced by the reference's own geodesics.py (tests/golden/reference_geodesics_golden.npz).

Same idea as make_reference_golden.py, one step further: /root/reference/mahakala/geodesics.py is executed UNMODIFIED
against a NumPy stand-in for the JAX names it imports:

    jit           identity                     vmap          Python loop over the mapped axis
    lax.scan      Python loop (a fixed point of the pure body is replicated instead of recomputed)
    lax.select    numpy.where                  inv           numpy.linalg.inv
    jacfwd(f)     complex-step derivative  d f / d x_k = Im f(x + i h e_k) / h,  h = 1e-30

The complex step evaluates the reference's own `metric` text on x + i h e_k; for an analytic expression it returns the
derivative to rounding error (no subtractive cancellation), which is what forward-mode autodiff returns as well, so
`rhs` below is the reference's contraction (geodesics.py:301-309) of the derivatives of the reference's metric
(:88-104) with the reference's LU inverse (:339-347).  Everything else (camera, nullification, RK4, the step rule,
freeze/reject logic, the +2 truncation, the last-point rule, the bisection) runs verbatim.  Differences to real JAX:
libm vs XLA transcendentals (only sqrt, cos, sin here: correctly rounded in both) and XLA's freedom to contract
a*b+c into FMAs, i.e. last-bit effects, amplified along trajectories as for any two implementations
(DESIGN.md noise floor).  /root/reference exists only in the build container: run here, commit the .npz.

    python tests/golden/make_reference_geodesics_golden.py        (about two minutes)
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_reference_golden import REF, _Arr, arr, load  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_geodesics_golden.npz")
H = 1e-30


def _wrap(fn):
    def g(*a, **k):
        r = fn(*a, **k)
        if isinstance(r, np.ndarray):
            return r.view(_Arr)
        if isinstance(r, (tuple, list)):
            return type(r)(x.view(_Arr) if isinstance(x, np.ndarray) else x for x in r)
        return r
    return g


def _jacfwd(f):
    def jac(x, *rest):
        x = np.asarray(x, dtype=np.float64)
        cols = []
        for k in range(x.shape[0]):
            xc = x.astype(np.complex128)
            xc[k] += 1j * H
            cols.append(np.asarray(f(xc.view(_Arr), *rest)).imag / H)
        return np.stack(cols, axis=-1).view(_Arr)
    return jac


def _vmap(f, in_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.asarray(a).shape[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            call = [(np.take(np.asarray(a), i, axis=ax).view(_Arr) if ax is not None else a) for a, ax in zip(args, axes)]
            outs.append(np.asarray(f(*call)))
        return np.stack(outs, axis=0).view(_Arr)
    return mapped


def _scan(f, init, xs):
    """lax.scan for a PURE body: once (carry, y) repeats, all later iterations are identical and are replicated"""
    carry = np.array(init, copy=True).view(_Arr)
    ys = None
    prev = None
    for it, x in enumerate(xs):
        new_carry, y = f(carry, x)
        if ys is None:
            ys = [np.empty((len(xs),) + np.asarray(c).shape, dtype=np.asarray(c).dtype) for c in y]
        for buf, c in zip(ys, y):
            buf[it] = c
        same = prev is not None and np.array_equal(new_carry, carry, equal_nan=True) and \
            all(np.array_equal(np.asarray(c), p, equal_nan=True) for c, p in zip(y, prev))
        carry, prev = np.array(new_carry, copy=True).view(_Arr), [np.array(c, copy=True) for c in y]
        if same:
            for buf in ys:
                buf[it + 1:] = buf[it]
            break
    return carry, tuple(b.view(_Arr) for b in ys)


def install_stand_in():
    jnp = types.ModuleType("jax.numpy")
    for name in ("select", "isclose", "heaviside", "sqrt", "minimum", "maximum", "concatenate", "diag", "array", "ones",
                 "zeros", "cross", "meshgrid", "linspace", "logical_or", "isnan", "abs", "zeros_like", "broadcast_to",
                 "all", "argmax", "arange", "multiply"):
        setattr(jnp, name, _wrap(getattr(np, name)))
    jnp.nan, jnp.newaxis, jnp.pi = np.nan, np.newaxis, np.pi
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.inv = _wrap(np.linalg.inv)
    jnp.linalg = linalg
    lax = types.ModuleType("jax.lax")
    lax.select = _wrap(lambda pred, a, b: np.where(pred, a, b))
    lax.scan = _scan
    jax = types.ModuleType("jax")
    jax.numpy, jax.lax = jnp, lax
    jax.jit = lambda f: f
    jax.jacfwd = _jacfwd
    jax.vmap = _vmap
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": linalg, "jax.lax": lax})
    pkg = types.ModuleType("mahakala")
    pkg.__path__ = [REF]
    sys.modules["mahakala"] = pkg


def compute():
    install_stand_in()
    geo = load("mahakala.geodesics", "geodesics.py")
    out = {}
    a = 0.94
    rng = np.random.default_rng(7)
    # metric / imetric / rhs at points from far away down to just outside the horizon (r_H = 1.34)
    pts = np.array([[0., 1000., 3., 500.], [0., 30., -20., 10.], [0., 5., 1., -2.], [0., 2.0, 0.5, 0.3],
                    [0., 1.2, -0.9, 0.4], [0., 0.3, 1.5, 0.1]])
    vel = np.concatenate([np.ones((6, 1)), rng.normal(0, 0.6, (6, 3))], axis=1)
    states = np.concatenate([pts, vel], axis=1)
    out["pt_states"] = states
    out["pt_metric"] = np.stack([np.asarray(geo.metric(arr(p), a)) for p in pts])
    out["pt_imetric"] = np.stack([np.asarray(geo.imetric(arr(p), a)) for p in pts])
    out["pt_rhs"] = np.stack([np.asarray(geo.rhs(arr(s), a)) for s in states])
    out["pt_radius"] = np.asarray(geo.radius_cal(arr(pts), a))
    out["pt_rk4_dt"] = -np.array([20., 0.7, 0.09, 0.016, 0.004, 0.005])
    out["pt_rk4"] = np.asarray(geo.RK4_gen(arr(states), arr(out["pt_rk4_dt"]), a))
    # cameras
    s0 = np.asarray(geo.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, 6))
    out["cam_grid_a094_i60_res6"] = s0
    out["cam_equator_a0_res8"] = np.asarray(geo.initialize_geodesics_at_camera(0.0, 60, 1000, -15, 15, 8, camera_type='Equator'))
    xr, vr = geo.get_camera_pixel(52, 1000, arr(np.array([3.0, 5.2, 7.5])), arr(np.array([0.3, 2.0, 4.4])))
    out["cam_pixel_x"], out["cam_pixel_v"] = np.asarray(xr), np.asarray(vr)
    # trajectories: the 36 rays of the 6x6 grid (captured and escaped rays), shadow-finder settings
    S, dt = geo.geodesic_integrator(2000, arr(s0), 40, 1e-2, a)
    out["traj_S"], out["traj_dt"] = np.asarray(S), np.asarray(dt)
    # a few rays with the imaging settings (tol 1e-4), N small enough that one ray hits the iteration cap
    S2, dt2 = geo.geodesic_integrator(450, arr(s0[[0, 14, 15, 21]]), 40, 1e-4, a)
    out["traj_cap_S"], out["traj_cap_dt"] = np.asarray(S2), np.asarray(dt2)
    # last-point rule and the bisection itself
    ang = np.array([0.0, 1.1, 2.2, 3.3, 4.4, 5.5])
    out["shadow_angles"] = ang
    out["select_r"] = np.asarray(geo.select_photons_integrator(60, ang, np.array([2.0, 4.0, 5.0, 5.5, 6.0, 8.0]), a))
    out["shadow_radii_a094_i60"] = np.asarray(geo.find_shadow_bisection_angles(a, 60, ang))
    return out


if __name__ == "__main__":
    res = compute()
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, {k: np.asarray(v).shape for k, v in res.items()})
