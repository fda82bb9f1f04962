"""The ARITHMETIC of the ray kernels, compiled for the host from the product's own headers (tests/host_harness), held to
the oracle in the CPU suite: closed-form Kerr-Schild acceleration (ks_metric.cuh), RK4 step and step rule
(integrate.cuh), the per-ray loop of integrate_kernel.cuh, and the fused emission chain (sample.cuh::emission_fast).

The build container has no GPU, so without this the kernels' numerics would only ever be checked at round end.  The
device differs from this build in one place only: the MUFU seeds of 1/x and 1/sqrt(x) (emulated here by library values
truncated to 22 bits); everything downstream is the same FMA sequence.  The -m gpu tests remain the parity tests
proper; nothing in the product loads this library (test_product_package_never_imports_the_oracle covers tests/ too).
"""
import ctypes

import numpy as np
import pytest

A = 0.94
_dp = ctypes.POINTER(ctypes.c_double)


def _d(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def hk(built):
    from host_harness import build
    return build.lib()


def _rhs(hk, P, which="hk_rhs"):
    out = np.empty_like(P)
    getattr(hk, which)(ctypes.c_long(P.shape[0]), _d(P), ctypes.c_double(A), _d(out))
    return out


def test_closed_form_acceleration_matches_the_oracle(hk):
    """KerrSchild::accel (expansion + twist form of grad l, 87 FP64 operations) and the round-1 component-wise form
    against the oracle's jets + 4x4 inverse (geodesics.py:294-309): 1e-12 at generic points, 1e-9 along trajectories
    that graze the horizon (where the oracle's own inverse loses digits)."""
    from oracle import c_oracle, mahakala_oracle as onp
    rng = np.random.default_rng(0)
    rnd = np.concatenate([np.zeros((6000, 1)), rng.normal(0, 6, (6000, 3)), rng.normal(0, 1, (6000, 4))], axis=1)
    rnd = np.ascontiguousarray(rnd[onp.radius_cal(rnd, A) > 1.4])
    s0 = np.ascontiguousarray(onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 16))
    S, dt = c_oracle.geodesic_integrator(10000, s0, 40, 1e-4, A)
    traj = S[::5].reshape(-1, 8)
    traj = np.ascontiguousarray(traj[onp.radius_cal(traj, A) > 1.36])
    for P, tol in ((rnd, 1e-12), (traj, 1e-9)):
        ref = c_oracle.rhs(P, A)
        scale = np.abs(ref[:, 4:]).max(axis=1, keepdims=True)
        new, old = _rhs(hk, P), _rhs(hk, P, "hk_rhs_v1")
        assert np.array_equal(new[:, :4], P[:, 4:])
        assert (np.abs(new[:, 4:] - ref[:, 4:]) / scale).max() < tol
        assert (np.abs(old[:, 4:] - ref[:, 4:]) / scale).max() < tol
        assert (np.abs(new[:, 4:] - old[:, 4:]) / scale).max() < tol


def test_rk4_step_matches_the_oracle(hk):
    from oracle import mahakala_oracle as onp
    s0 = np.ascontiguousarray(onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 12))
    rng = np.random.default_rng(1)
    pts = s0.copy()
    pts[:, 1:4] *= rng.uniform(0.004, 0.05, (pts.shape[0], 1))          # r from 4 to 50 along the camera rays
    dt = -(onp.radius_cal(pts, A) - onp.radius_EH(A)) / 40.
    out = np.empty_like(pts)
    hk.hk_rk4(ctypes.c_long(pts.shape[0]), _d(pts), _d(np.ascontiguousarray(dt)), ctypes.c_double(A), _d(out))
    ref = onp.RK4_gen(pts, dt, A)
    assert (np.abs(out - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() < 1e-13


@pytest.mark.parametrize("res,N,tol,sub", [(64, 2000, 1e-2, 1), (1024, 10000, 1e-4, 16)])
def test_ray_loop_matches_the_oracle(hk, res, N, tol, sub):
    """BASELINE config 1 (full 64x64 grid) and the every-16th-pixel sub-lattice of config 2 through the per-ray loop of
    the integrate kernel: classification bit-exact, step counts identical on escaped rays, end states 1e-9."""
    from oracle import c_oracle, mahakala_oracle as onp
    s0 = np.ascontiguousarray(onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, res))
    if sub > 1:
        idx = (np.arange(0, res, sub)[:, None] * res + np.arange(0, res, sub)[None, :]).reshape(-1)
        s0 = np.ascontiguousarray(s0[idx])
    n = s0.shape[0]
    fin, ns, rl = np.empty((n, 8)), np.empty(n, dtype=np.int32), np.empty(n)
    hk.hk_integrate(ctypes.c_long(n), _d(s0), ctypes.c_int(N), ctypes.c_double(40.), ctypes.c_double(tol),
                    ctypes.c_double(A), _d(fin), ns.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _d(rl))
    ref = c_oracle.integrate(N, s0, 40, tol, A)
    cap = ref["r_last"] < 100
    assert np.array_equal(rl < 100, cap) and cap.sum() in (792, 799)
    esc = ~cap
    assert np.array_equal(ns[esc], ref["nsteps"][esc])
    assert abs(int(ns.sum()) - int(ref["nsteps"].sum())) <= 4 * cap.sum()
    err = np.abs(fin[esc] - ref["final"][esc]).max(axis=1) / np.abs(ref["final"][esc]).max(axis=1)
    assert err.max() < 1e-9
    assert np.allclose(rl[esc], ref["r_last"][esc], rtol=1e-9)


def _adaptive(hk, s0, rtol, tol=1e-2, N=100000, cap=0.5, atol=1e-12):
    n = s0.shape[0]
    fin, ns, nr, rl = np.empty((n, 8)), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32), np.empty(n)
    ip = ctypes.POINTER(ctypes.c_int)
    hk.hk_integrate_adaptive(ctypes.c_long(n), _d(s0), ctypes.c_int(N), ctypes.c_double(rtol), ctypes.c_double(atol),
                             ctypes.c_double(tol), ctypes.c_double(cap), ctypes.c_double(A), _d(fin),
                             ns.ctypes.data_as(ip), nr.ctypes.data_as(ip), _d(rl))
    return fin, ns, nr, rl


def constants_of_motion(S, a):
    """energy -k_t, axial angular momentum k_phi = x k_y - y k_x (covariant components) and the norm g(k, k)"""
    from oracle import mahakala_oracle as onp
    g = np.stack([onp.metric(x[:4], a) for x in S])
    kcov = np.einsum('nij,nj->ni', g, S[:, 4:])
    return -kcov[:, 0], S[:, 1] * kcov[:, 2] - S[:, 2] * kcov[:, 1], np.einsum('ni,ni->n', kcov, S[:, 4:])


def test_adaptive_integrator_option(hk):
    """The optional embedded Dormand-Prince 5(4) integrator (csrc/adaptive.cuh; not in the reference)
    on BASELINE config 1: same captured / escaped classification as the reference's fixed-rule RK4 (792 captured),
    constants of motion conserved in proportion to rtol and better than the fixed rule does with twice the
    acceleration evaluations, no rejected steps at the default tolerance, frozen states inside the live range, the
    step cap N."""
    from oracle import c_oracle, mahakala_oracle as onp
    s0 = np.ascontiguousarray(onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 64))
    ref = c_oracle.integrate(2000, s0, 40, 1e-2, A)
    cap_ref = ref["r_last"] < 100
    E0, L0, _ = constants_of_motion(s0, A)
    rH = onp.radius_EH(A)
    worst = {}
    for rtol in (1e-6, 1e-9, 1e-11):
        fin, ns, nr, rl = _adaptive(hk, s0, rtol)
        cap = rl < 100
        assert np.array_equal(cap, cap_ref) and cap.sum() == 792
        assert np.isfinite(fin).all()
        assert np.allclose(rl, onp.radius_cal(fin, A), rtol=1e-12)
        assert ((rl - rH >= 1e-2) & (rl - rH <= 1500)).all()               # frozen inside the live range
        assert (rl[cap] - rH).max() < 2e-2 and rl[~cap].min() > 500        # ... right at its two ends
        E1, L1, n1 = constants_of_motion(fin, A)
        worst[rtol] = max(np.abs(E1 / E0 - 1)[~cap].max(), (np.abs(L1 - L0) / np.abs(L0).max())[~cap].max(),
                          np.abs(n1)[~cap].max())
        if rtol <= 1e-9:
            assert nr.sum() == 0
        if rtol == 1e-9:
            assert 6 * ns.sum() < 0.6 * 4 * ref["nsteps"].sum()            # fewer acceleration evaluations than RK4 ...
            assert np.abs(E1 / E0 - 1)[cap].max() < 1e-5                   # ... and captured rays stay under control
    Er, Lr, nrm = constants_of_motion(ref["final"], A)
    rk4 = max(np.abs(Er / E0 - 1)[~cap_ref].max(), (np.abs(Lr - L0) / np.abs(L0).max())[~cap_ref].max(),
              np.abs(nrm)[~cap_ref].max())
    assert worst[1e-6] < 1e-3 and worst[1e-9] < 1e-7 and worst[1e-11] < 1e-9 and worst[1e-9] < rk4
    # step cap, and a ray that starts outside the live range
    fin, ns, nr, rl = _adaptive(hk, s0[:64], 1e-9, N=7)
    assert ns.max() == 7
    far = s0[:4].copy(); far[:, 1:4] *= 3.0                                # r = 3000 M > r_H + 1500
    fin, ns, nr, rl = _adaptive(hk, np.ascontiguousarray(far), 1e-9)
    assert (ns == 0).all() and np.array_equal(fin, far)


def test_adaptive_integrator_against_an_independent_solver(hk):
    """Accuracy pin of the adaptive option by a solver that shares nothing with it: SciPy's DOP853 (8th order, rtol
    1e-12) on the ORACLE's literal right-hand side (jets + 4x4 inverse, geodesics.py:294-309), stopped by an event at the
    coordinate time at which the adaptive run (rtol 1e-10, closed-form acceleration) froze.  Escaped rays agree to 1e-9 of
    the state, captured ones -- whose wavevector blue-shifts by orders of magnitude near the horizon -- to 1e-7."""
    from scipy.integrate import solve_ivp
    from oracle import c_oracle, mahakala_oracle as onp
    s0 = np.ascontiguousarray(onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 64))
    idx = [0, 700, 1500, 2079, 2080, 2111, 3000, 4095]
    sub = np.ascontiguousarray(s0[idx])
    fin, ns, nr, rl = _adaptive(hk, sub, 1e-10)
    captured = rl < 100
    assert captured.sum() == 3 and nr.sum() == 0

    def f(lam, y):
        return c_oracle.rhs(y[None, :].copy(), A)[0]

    for k in range(len(idx)):
        def reached(lam, y, t_end=fin[k, 0]):
            return y[0] - t_end
        reached.terminal = True
        sol = solve_ivp(f, (0.0, -1e6), sub[k], method="DOP853", rtol=1e-12, atol=1e-14, events=reached, max_step=50.0)
        assert sol.status == 1 and len(sol.y_events[0]) == 1
        ye = sol.y_events[0][0]
        err = np.abs(ye - fin[k]) / np.maximum(np.abs(fin[k]), 1e-3 * np.abs(fin[k]).max())
        assert err.max() < (1e-7 if captured[k] else 1e-9), (idx[k], err.max())


def test_fused_emission_chain_on_adversarial_inputs(hk, monkeypatch):
    """The adversarial cases of tests/test_emission_gpu.py (sigma cut, Theta_e floor, X range, Planck switch, aligned /
    reversed / null wavevectors, degenerate primitives, exp underflow) run through the HOST build of emission_fast."""
    import test_emission_gpu as T
    from mahakala_b200 import transfer

    def host_probe(S, pc, a, P, nus, fast=True):
        S, pc = np.ascontiguousarray(S), np.ascontiguousarray(np.asarray(pc))
        nus = np.ascontiguousarray(np.atleast_1d(nus), dtype=np.float64)
        n = S.shape[0]
        em, ab = np.empty((nus.size, n)), np.empty((nus.size, n))
        pv = np.array([getattr(P, k) for k, _ in P._fields_])
        hk.hk_emission_fast(ctypes.c_long(n), _d(S), _d(pc), _d(pv), ctypes.c_double(a), ctypes.c_int(nus.size),
                            _d(nus), _d(em), _d(ab))
        return em, ab

    monkeypatch.setattr(transfer, "emission_probe", host_probe)
    for name in ("test_sigma_cut_both_sides", "test_theta_floor_both_sides", "test_x_range_and_cube_root_fallback",
                 "test_planck_series_switch", "test_field_aligned_reversed_and_null_wavevectors",
                 "test_degenerate_primitives", "test_exp_underflow_band"):
        getattr(T, name)(True)


def _host_snapshot(om, storage=np.float64):
    """Host arrays in the device snapshot's layout from an oracle model (reference layout all_meshblocks)."""
    from mahakala_b200.grmhd.athenak import build_block_grid
    amb = om.all_meshblocks                                   # (nmb, 8, nk+2, nj+2, ni+2), file order
    canon = [0, 4, 1, 2, 3, 5, 6, 7]                          # dens, eint, U1..3, B1..3
    cells = np.ascontiguousarray(amb[:, canon].transpose(0, 2, 3, 4, 1)).astype(storage)
    nmb = amb.shape[0]
    dx = [np.asarray(v)[:, 1] - np.asarray(v)[:, 0] for v in (om.x1v, om.x2v, om.x3v)]
    geom = np.zeros((nmb, 16))
    geom[:, 0:3] = np.stack([om.x1f[:, 0], om.x2f[:, 0], om.x3f[:, 0]], axis=1)
    geom[:, 3:6] = np.stack([om.x1f[:, -1], om.x2f[:, -1], om.x3f[:, -1]], axis=1)
    geom[:, 6:9] = np.stack([om.x1v[:, 0], om.x2v[:, 0], om.x3v[:, 0]], axis=1)
    geom[:, 9:12] = np.stack(dx, axis=1)
    geom[:, 12:15] = 1.0 / geom[:, 9:12]
    pow2 = bool(np.all(np.frexp(geom[:, 9:12])[0] == 0.5))
    grid, gn, g0, ginv = build_block_grid(om.x1f, om.x2f, om.x3f)
    lo, hi = geom[:, 0:3].min(axis=0), geom[:, 3:6].max(axis=0)
    return dict(cells=cells, geom=geom, grid=np.ascontiguousarray(grid), gn=gn, g0=g0, ginv=ginv, lo=lo, hi=hi,
                pow2=pow2, shape=amb.shape)


def _host_sample(hk, snap, S, kind):
    S = np.ascontiguousarray(S.reshape(-1, 8))
    out = np.empty((8, S.shape[0]))
    nmb, _, nk2, nj2, ni2 = snap["shape"]
    ip = ctypes.POINTER(ctypes.c_int)
    hk.hk_sample_prims(ctypes.c_int(kind), snap["cells"].ctypes.data_as(ctypes.c_void_p),
                       ctypes.c_int(1 if snap["cells"].dtype == np.float32 else 0), ctypes.c_int(nmb),
                       ctypes.c_int(nk2 - 2), ctypes.c_int(nj2 - 2), ctypes.c_int(ni2 - 2), _d(snap["geom"]),
                       snap["grid"].ctypes.data_as(ip), snap["gn"].ctypes.data_as(ip), _d(snap["g0"]), _d(snap["ginv"]),
                       _d(snap["lo"]), _d(snap["hi"]), ctypes.c_int(1 if snap["pow2"] else 0),
                       ctypes.c_long(S.shape[0]), _d(S), _d(out))
    return out


def test_sampling_path_matches_the_oracle(hk):
    """Block lookup + cell index + trilinear gather of sample.cuh (generic path and the two specialisations of the
    fused kernel, whose cell index multiplies by the exact 1/dx of power-of-two meshes) against the oracle's literal
    athenak.py:663-757 on trajectories, on points exactly on faces / corners, outside, and NaN."""
    from helpers import oracle_model, snapshot_arrays
    from oracle import c_oracle, mahakala_oracle as onp
    om = oracle_model(snapshot_arrays(ncells=32, block=16, extent=16.0, funnel={}), A)
    s0 = np.ascontiguousarray(onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 10))
    S, dt = c_oracle.geodesic_integrator(10000, s0, 40, 1e-4, A)
    f = om.x1f[0]
    edge = [[0.0, x, y, z, 1.0, 0.3, -0.2, 0.1]
            for x in (f[0], f[1], f[-1], -16.0, 16.0, 0.0, 15.999999999999998, 16.000000000000004, -15.75, 1e300, np.nan)
            for y in (0.0, -16.0, 16.0, 2.0) for z in (0.5, 0.0, 16.0, -16.0)]
    pts = np.concatenate([S.reshape(-1, 8), np.array(edge)])
    ref = c_oracle.sample(om, pts, mode="prims")
    ref = np.stack([ref[k] for k in ("dens", "u", "U1", "U2", "U3", "B1", "B2", "B3")])
    assert (ref[0] > 0).sum() > 5000 and (ref[0] == 0).sum() > 5000
    snap64, snap32 = _host_snapshot(om), _host_snapshot(om, np.float32)
    assert snap64["pow2"]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    results = {}
    for name, snap, kind in (("generic", snap64, 0), ("f64 pow2", snap64, 1), ("f32 pow2", snap32, 2),
                             ("generic f32", snap32, 0)):
        got = _host_sample(hk, snap, pts, kind)
        assert np.array_equal(got == 0, ref == 0), name
        assert (np.abs(got - ref) / scale).max() < 1e-14, name
        results[name] = got
    # the specialisations return the very same numbers as the generic path (values are float32-representable)
    for name in ("f64 pow2", "f32 pow2", "generic f32"):
        assert np.array_equal(results[name], results["generic"]), name
    # a mesh whose cell size is NOT a power of two keeps the corrected division: still the oracle's numbers
    arr = snapshot_arrays(ncells=24, block=12, extent=15.0)
    om2 = oracle_model(arr, A)
    snap = _host_snapshot(om2)
    assert not snap["pow2"]
    pts2 = np.ascontiguousarray(S[150:400].reshape(-1, 8))
    ref2 = c_oracle.sample(om2, pts2, mode="prims")
    ref2 = np.stack([ref2[k] for k in ("dens", "u", "U1", "U2", "U3", "B1", "B2", "B3")])
    got2 = _host_sample(hk, snap, pts2, 0)
    assert (ref2[0] > 0).sum() > 2000 and np.array_equal(got2 == 0, ref2 == 0)
    assert (np.abs(got2 - ref2) / np.abs(ref2).max(axis=1, keepdims=True)).max() < 1e-14
