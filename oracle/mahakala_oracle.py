"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

A literal float64 NumPy restatement of the reference's per-ray hot path (camera-ray initialisation,
RK4 Kerr-Schild geodesics, AthenaK-style sampling, thermal-synchrotron transfer).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; the product
package ``mahakala_b200`` never does (and fails loudly when its CUDA library is missing).

Parity pinning: the reference (JAX) cannot be imported in this image (no jax / jaxlib / h5py, no
network), so this oracle is pinned against

* the reference's only golden vectors, ``tests/data/shadow_data.npy`` (4 cases; re-saved pickle-free as
  ``tests/golden/shadow_golden.npz``), through ``find_shadow_bisection_angles`` — this pins camera
  (polar) + nullify + RK4 integrator + classifier; checked in ``tests/test_oracle_golden.py``;
* a line-for-line ``torch.func.jacfwd``/``vmap``/``linalg.inv`` transliteration of ``geodesics.py:294-347``
  (``oracle/literal_torch.py``) and 50-digit mpmath evaluation of the metric derivatives.

* OUTPUTS OF THE REFERENCE'S OWN SOURCE FILES run in the build container against a NumPy stand-in for the JAX
  names they import (``tests/golden/make_reference_golden.py``: constants.py, electrons.py, grmhd/grmhd.py,
  transfer.py; ``tests/golden/make_reference_geodesics_golden.py``: the whole of geodesics.py, with ``jacfwd``
  replaced by a complex-step derivative of the reference's metric).  Frozen in ``tests/golden/reference_golden.npz``
  and ``reference_geodesics_golden.npz`` and checked in ``tests/test_oracle_cpu.py``: metric, inverse, radius and the
  three cameras agree to the last bit or two, rhs to 1e-15, stored trajectories have the identical shape, freeze
  pattern and step counts with end states at 3e-15 (escaped) / 9e-14 (captured), the shadow radii and the transfer
  scans are bit-identical, Theta_e / j / alpha agree to 1e-13 with identical zero / NaN patterns.

* OUTPUTS OF THE WHOLE REFERENCE PACKAGE (real ``__init__``, athenak.py loader and sampling, images.make_image)
  imported with the same stand-in plus an in-memory ``h5py`` (``tests/golden/make_reference_fluid_golden.py`` ->
  ``reference_fluid_golden.npz``): ghost-zone fill bit-identical on a single-level snapshot (and different from the
  reference only in the edge ghost cells its AMR branch gets wrong), sampled primitives bit-identical, fluid scalars
  1e-16, images 1e-13 per pixel.

What no vector covers is XLA itself (its libm and optional FMA contraction): the stand-in evaluates the reference's
text under NumPy's IEEE double arithmetic.  Each function cites the reference file:line (relative to
/root/reference/mahakala/) it follows.

Forward-mode differentiation (``jacfwd(metric)``, geodesics.py:305) is restated with an explicit jet
(value + 4 tangents) pushed through the *same* metric expression, the matrix inverse (geodesics.py:347)
with ``numpy.linalg.inv`` (LAPACK getrf/getri, the same algorithm family XLA's CPU backend calls).
"""
import numpy as np

# ----------------------------------------------------------------------------------------------
# constants — constants.py:23-31, electrons.py:23-29, grmhd/grmhd.py:26-32 (digit for digit)
# ----------------------------------------------------------------------------------------------
EE = 4.8032e-10
KB = 1.3807e-16
CL = 2.99792458e10
ME = 9.1094e-28
MP = 1.6726e-24
EC = 4.8032e-10
HPL = 6.6261e-27
GNEWT = 6.6743e-8
Msun = 1.989e33


# ----------------------------------------------------------------------------------------------
# jets: value + tangents, the literal meaning of jacfwd
# ----------------------------------------------------------------------------------------------
class _Jet:
    """value ``v`` of shape (n,), tangents ``d`` of shape (n, 4) = d/dx^k."""
    __slots__ = ("v", "d")
    __array_priority__ = 1000

    def __init__(self, v, d):
        self.v = v
        self.d = d

    @staticmethod
    def const(c, like):
        return _Jet(np.zeros_like(like.v) + c, np.zeros_like(like.d))

    def _lift(self, o):
        return o if isinstance(o, _Jet) else _Jet.const(o, self)

    def __add__(self, o):
        o = self._lift(o)
        return _Jet(self.v + o.v, self.d + o.d)
    __radd__ = __add__

    def __sub__(self, o):
        o = self._lift(o)
        return _Jet(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        o = self._lift(o)
        return _Jet(self.v * o.v, self.d * o.v[:, None] + o.d * self.v[:, None])
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._lift(o)
        q = self.v / o.v
        return _Jet(q, (self.d - o.d * q[:, None]) / o.v[:, None])

    def __rtruediv__(self, o):
        return self._lift(o) / self

    def sqrt(self):
        s = np.sqrt(self.v)
        return _Jet(s, self.d / (2.0 * s)[:, None])


def _sqrt(x):
    return x.sqrt() if isinstance(x, _Jet) else np.sqrt(x)


def _metric_parts(x1, x2, x3, a):
    """The body of ``metric`` (geodesics.py:95-103), generic over float arrays and jets."""
    aa = a * a
    zz = x3 * x3                      # x[3]**2.
    kk = 0.5 * (x1 * x1 + x2 * x2 + zz - aa)
    rr = _sqrt(kk * kk + aa * zz) + kk
    r = _sqrt(rr)
    f = (2.0 * rr * r) / (rr * rr + aa * zz)
    l1 = (r * x1 + a * x2) / (rr + aa)
    l2 = (r * x2 - a * x1) / (rr + aa)
    l3 = x3 / r
    return f, l1, l2, l3


def metric(x, bhspin):
    """geodesics.py:88-104.  x: (..., 4) -> g: (..., 4, 4)."""
    x = np.asarray(x, dtype=np.float64)
    f, l1, l2, l3 = _metric_parts(x[..., 1], x[..., 2], x[..., 3], bhspin)
    l = np.stack([np.ones_like(f), l1, l2, l3], axis=-1)
    eta = np.diag([-1.0, 1.0, 1.0, 1.0])
    return eta + f[..., None, None] * (l[..., :, None] * l[..., None, :])


def imetric(x, bhspin):
    """geodesics.py:339-347: numerical inverse of the covariant metric."""
    return np.linalg.inv(metric(x, bhspin))


def metric_jac(x, bhspin):
    """``jacfwd(metric)(x, bhspin)`` (geodesics.py:305): jg[..., i, j, k] = d g_ij / d x^k.  x: (n, 4)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    eye = np.eye(4)
    jets = [_Jet(x[:, m].copy(), np.broadcast_to(eye[m], (n, 4)).copy()) for m in range(4)]
    f, l1, l2, l3 = _metric_parts(jets[1], jets[2], jets[3], bhspin)
    l = [_Jet.const(1.0, f), l1, l2, l3]
    jg = np.zeros((n, 4, 4, 4))
    for i in range(4):
        for j in range(4):
            gij = f * (l[i] * l[j])          # eta is constant
            jg[:, i, j, :] = gij.d
    return jg


def rhs(state, bhspin):
    """geodesics.py:294-309 (vectorised as :312-314).  state: (n, 8) -> (n, 8)."""
    x = state[:, :4]
    v = state[:, 4:]
    ig = imetric(x, bhspin)
    jg = metric_jac(x, bhspin)
    t1 = np.einsum('nij,nj->ni', np.einsum('nijk,nk->nij', jg, v), v)     # (jg @ v) @ v
    t2 = np.einsum('ni,nik->nk', v, np.einsum('nj,nijk->nik', v, jg))     # v @ (v @ jg)
    a = np.einsum('nij,nj->ni', ig, -t1 + 0.5 * t2)
    return np.concatenate([v, a], axis=1)


def RK4_gen(state1, dt, bhspin):
    """geodesics.py:317-336."""
    val = len(state1)
    ans1 = rhs(state1, bhspin)
    k1 = np.multiply(dt.reshape(val, 1), ans1)
    ans1 = rhs(state1 + 0.5 * k1, bhspin)
    k2 = np.multiply(dt.reshape(val, 1), ans1)
    ans1 = rhs(state1 + 0.5 * k2, bhspin)
    k3 = np.multiply(dt.reshape(val, 1), ans1)
    ans1 = rhs(state1 + k3, bhspin)
    k4 = np.multiply(dt.reshape(val, 1), ans1)
    return state1 + 1 / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


def radius_cal(x, bhspin):
    """geodesics.py:284-291."""
    x = np.asarray(x)
    R = np.sqrt(x[..., 1]**2 + x[..., 2]**2 + x[..., 3]**2)
    return np.sqrt((R**2 - bhspin**2 + np.sqrt((R**2 - bhspin**2)**2 + 4 * bhspin**2 * x[..., 3]**2)) / 2)


def radius_EH(a_spin):
    """geodesics.py:350-351."""
    return 1 + np.sqrt(1 - a_spin**2)


def _step_rule(state, div, tol, bhspin):
    """geodesics.py:249-252 / :258-261."""
    with np.errstate(invalid='ignore'):
        dt = -(radius_cal(state[:, :4], bhspin) - radius_EH(bhspin)) / div
        cond = np.logical_or(np.isnan(dt), np.abs(dt) * div < tol)
        cond = np.logical_or(cond, np.abs(dt) * div > 1500)
    return np.where(cond, 0.0, dt)


def geodesic_integrator(N, s0, div, tol, bhspin):
    """geodesics.py:233-281.

    ``lax.scan`` runs all N iterations; a frozen ray re-proposes the same rejected step forever
    (deterministic), so we only advance rays that are still moving and stop at the first all-zero row,
    then reproduce the truncation ``[:first_zero_idx + 2]`` (:275-281) exactly.
    """
    s = np.array(s0, dtype=np.float64)
    npx = s.shape[0]
    states, dts = [], []
    active = np.ones(npx, dtype=bool)
    first_zero_idx = None
    with np.errstate(all='ignore'):
        for it in range(N):
            dt_row = np.zeros(npx)
            idx = np.nonzero(active)[0]
            if idx.size:
                cur = s[idx]
                dt = _step_rule(cur, div, tol, bhspin)
                new_state = RK4_gen(cur, dt, bhspin)
                dtnew = _step_rule(new_state, div, tol, bhspin)
                rej = dtnew == 0.
                dt = np.where(rej, 0.0, dt)
                new_state = np.where(rej[:, None], cur, new_state)
                dt_row[idx] = dt
                states.append(s.copy())
                dts.append(dt_row)
                s[idx] = new_state
                active[idx[rej]] = False
            else:
                states.append(s.copy())
                dts.append(dt_row)
            if not np.any(dt_row != 0):
                first_zero_idx = it
                break
    if first_zero_idx is None or first_zero_idx < 1:
        first_zero_idx = N
    first_zero_idx += 2
    nrows = min(first_zero_idx, N)
    while len(states) < nrows:          # rows after the first all-zero row repeat the frozen state
        states.append(s.copy())
        dts.append(np.zeros(npx))
    return np.stack(states[:nrows]), np.stack(dts[:nrows])


# ----------------------------------------------------------------------------------------------
# camera — geodesics.py:29-230
# ----------------------------------------------------------------------------------------------
def _Image_to_BH(x, y, z, i, d):
    """geodesics.py:204-209."""
    i = i * np.pi / 180
    x_BH = -y * np.cos(i) + z * np.sin(i) + d * np.sin(i)
    y_BH = x
    z_BH = y * np.sin(i) + z * np.cos(i) + d * np.cos(i)
    return np.array([x_BH, y_BH, z_BH])


def _perpendicular(a):
    """geodesics.py:212-216."""
    b = np.zeros_like(a)
    b[0] = a[0] + a[1]
    b[1] = a[1] - a[0]
    return b


def get_camera_pixel(inclination, distance, radius, angle):
    """geodesics.py:107-134."""
    size = np.size(radius)
    x = np.ones(size) * np.cos(angle) * radius
    y = np.ones(size) * np.sin(angle) * radius
    z = np.ones(size) * 0.
    origin_BH = _Image_to_BH(0, 0, 0, inclination, distance)
    temp_coord = _perpendicular(np.array([x, y]))
    init_BH = _Image_to_BH(x, y, z, inclination, distance)
    perp_BH = _Image_to_BH(temp_coord[0], temp_coord[1], np.zeros(1), inclination, distance)
    vec1 = - init_BH.T + origin_BH
    vec2 = perp_BH - init_BH
    k_vec = np.cross(vec1, vec2.T)
    s0_x = np.array([np.zeros(size), init_BH[0].flatten(), init_BH[1].flatten(), init_BH[2].flatten()])
    s0_v = np.array([np.ones(size), k_vec.T[0].flatten(), k_vec.T[1].flatten(), k_vec.T[2].flatten()])
    return s0_x, s0_v


def get_initial_grid(inclination, distance, fov_lower, fov_upper, spacing, camera_type):
    """geodesics.py:137-201."""
    if camera_type.lower() == 'grid':
        grid_list = np.linspace(fov_lower, fov_upper, 2 * spacing + 1)[1::2]
        n2 = len(grid_list)**2
        z = 0 * np.ones(n2)
        x, y = np.meshgrid(grid_list, grid_list, indexing='ij')
        x = x.flatten()
        y = y.flatten()
        origin_BH = _Image_to_BH(0, 0, 0, inclination, distance)
        temp_coord = _perpendicular(np.array([x, y]))
        init_BH = _Image_to_BH(x, y, z, inclination, distance)
        perp_BH = _Image_to_BH(temp_coord[0], temp_coord[1], 0 * np.ones(n2), inclination, distance)
        vec1 = - init_BH.T + origin_BH
        vec2 = perp_BH - init_BH
        k_vec = np.cross(vec1, vec2.T)
        s0_x = np.array([np.zeros(n2), init_BH[0].flatten(), init_BH[1].flatten(), init_BH[2].flatten()])
        s0_v = np.array([np.ones(n2), k_vec.T[0].flatten(), k_vec.T[1].flatten(), k_vec.T[2].flatten()])
        return s0_x, s0_v
    elif camera_type.lower() == 'equator':
        grid_list = np.linspace(fov_lower, fov_upper, 2 * spacing + 1)[1::2]
        s0_x = np.zeros((4, len(grid_list)))
        s0_x[1] = distance
        s0_x[2] = grid_list
        s0_v = np.ones((4, len(grid_list)))
        s0_v[2] = 0
        s0_v[3] = 0
        return s0_x, s0_v
    else:
        print(f'Unexpected camera type "{camera_type}".')
        print('Please choose either "grid" or "equator"')


def _quadratic(A, b, C):
    """geodesics.py:58-66 (isclose defaults rtol=1e-5, atol=1e-8; heaviside(b, 1))."""
    bb = b * b
    AC = A * C
    dd = np.where(~np.isclose(bb, AC), bb - AC, 0.0)
    bs = np.heaviside(b, 1)
    D = - (b + bs * np.sqrt(dd))
    x1 = D / A
    x2 = C / D
    return np.minimum(x1, x2), np.maximum(x1, x2)


def initial_condition(s0_x, s0_v, bhspin):
    """geodesics.py:219-230 with _Nullify (p=1) :69-85, vectorised over columns."""
    xs = np.asarray(s0_x, dtype=np.float64).T      # (n, 4)
    vs = np.asarray(s0_v, dtype=np.float64).T
    g = metric(xs, bhspin)
    with np.errstate(all='ignore'):
        A = vs[:, 0] * g[:, 0, 0] * vs[:, 0]
        b = np.einsum('ni,ni->n', vs[:, 1:], g[:, 1:, 0]) * vs[:, 0]
        C = np.einsum('ni,nij,nj->n', vs[:, 1:], g[:, 1:, 1:], vs[:, 1:])
        d1, d2 = _quadratic(A, b, C)
        S = np.where(d1 > 0, d1, np.where(d2 > 0, d2, np.nan))
        vnull = np.concatenate([vs[:, :1], vs[:, 1:] / S[:, None]], axis=1)
    return np.concatenate([xs, vnull], axis=1)


def initialize_geodesics_at_camera(bhspin, inclination, distance, fov_lower, fov_upper,
                                   pixels_per_side, camera_type='grid'):
    """geodesics.py:29-55."""
    s0_x, s0_v = get_initial_grid(inclination, distance, fov_lower, fov_upper, pixels_per_side,
                                  camera_type)
    return initial_condition(s0_x, s0_v, bhspin)


# ----------------------------------------------------------------------------------------------
# shadow finder — geodesics.py:354-435
# ----------------------------------------------------------------------------------------------
def last_point_radius(S, dt, bhspin):
    """geodesics.py:370-378 (the two ``.at[].set()`` results are discarded there, so they are no-ops)."""
    r = radius_cal(S, bhspin)
    maxi = np.argmax(dt, axis=0)
    maxi = maxi - 1                      # negative index wraps to the last row, as in the reference
    return r[maxi, np.arange(r.shape[1])]


def select_photons_integrator(inc, angle, radius, bhspin, distance=1000, max_steps=2000,
                              integrator=None):
    """geodesics.py:354-378.  ``integrator`` lets tests swap in the C oracle for speed."""
    s0_x, s0_v = get_camera_pixel(inc, distance, radius, angle)
    init_one = initial_condition(s0_x, s0_v, bhspin)
    integ = geodesic_integrator if integrator is None else integrator
    S, dt = integ(max_steps, init_one, 40, 1e-2, bhspin)
    return last_point_radius(S, dt, bhspin)


def find_shadow_bisection_angles(bhspin, inc, angles, max_steps=2000, error_allowed=0.001, max_it=40,
                                 integrator=None):
    """geodesics.py:405-435."""
    inner = np.zeros_like(angles) + 0.5
    outer = np.zeros_like(angles) + 10
    error = outer - inner
    bisection_limit = 100
    counter = 0
    while np.max(error) > error_allowed and counter < max_it:
        final_mid = select_photons_integrator(inc, angles, (outer - inner) / 2 + inner, bhspin,
                                              max_steps=max_steps, integrator=integrator)
        fell = np.where(final_mid < bisection_limit)
        got_away = np.where(final_mid >= bisection_limit)
        inner[fell] = (outer[fell] - inner[fell]) / 2 + inner[fell]
        outer[got_away] = (outer[got_away] - inner[got_away]) / 2 + inner[got_away]
        error = outer - inner
        counter += 1
    return inner


def find_shadow_bisection(bhspin, inc, num_angles, max_steps=2000, error_allowed=0.001, max_it=40,
                          integrator=None):
    """geodesics.py:381-402."""
    angles = np.arange(num_angles) / num_angles * 2. * np.pi
    radii = find_shadow_bisection_angles(bhspin, inc, angles, max_it=max_it,
                                         error_allowed=error_allowed, max_steps=max_steps,
                                         integrator=integrator)
    radii = np.append(radii, radii[0])
    angles = np.append(angles, angles[0])
    return angles, radii


# ----------------------------------------------------------------------------------------------
# electrons / transfer — electrons.py:32-50, transfer.py:30-144
# ----------------------------------------------------------------------------------------------
def rlow_rhigh_model(dens, u, beta, r_low=1, r_high=40, electron_gamma=4. / 3, ion_gamma=5. / 3):
    """electrons.py:32-50."""
    with np.errstate(all='ignore'):
        T_ratio = (r_high * beta**2 + r_low) / (1 + beta**2)
        t_electron = CL**2 * (MP * u * (electron_gamma - 1.) * (ion_gamma - 1.))
        t_electron = t_electron / (dens * ((ion_gamma - 1.) + (electron_gamma - 1.) * T_ratio))
        return t_electron / (ME * CL * CL)


def synchrotron_coefficients(Ne, Theta_e, B, pitch_angle, nu, invariant=True, rescale_nu=1.):
    """transfer.py:30-86."""
    nu_ratio_limit = 1.e12
    Theta_e_min = 0.3
    with np.errstate(all='ignore'):
        nuc = EE * B / (2. * np.pi * ME * CL)
        nus = (2. / 9.) * nuc * Theta_e**2 * np.sin(pitch_angle)
        X = nu / nus
        var = np.exp(- X**(1. / 3))
        term = np.sqrt(X) + 2.0**(11. / 12) * X**(1. / 6)
        emissivity = Ne * nus * term**2 / (2. * Theta_e**2.)
        emissivity = emissivity * var * np.sqrt(2) * np.pi * EE**2 / (3.0 * CL)
        emissivity = np.where(X > nu_ratio_limit, 0.0, emissivity)
        emissivity = np.where(Theta_e < Theta_e_min, 0.0, emissivity)
        bx = HPL * nu / (ME * CL * CL * Theta_e)
        series_expansion = bx / 24. * (24. + bx * (12. + bx * (4. + bx)))
        B_denominator = np.where(bx < 2.e-3, series_expansion, np.exp(bx) - 1)
        B_nu = (2. * HPL * nu**3. / B_denominator) / CL**2.
        absorptivity = emissivity / B_nu
        if invariant:
            rescaled_nu = nu * rescale_nu
            emissivity = emissivity / rescaled_nu**2.
            absorptivity = absorptivity * rescaled_nu
        emissivity = np.where(np.isnan(emissivity), 0.0, emissivity)
        absorptivity = np.where(np.isnan(absorptivity), 0.0, absorptivity)
    return emissivity, absorptivity


def solve_specific_intensity(emissivity, absorptivity, dt, L_unit, dIs=False):
    """transfer.py:89-119 (back-to-front explicit Euler)."""
    nsteps, npx = emissivity.shape
    I_nu = np.zeros(npx)
    out = []
    for i in range(nsteps - 1, 0, -1):
        dI = - dt[i - 1, :] * L_unit * (emissivity[i, :] - (absorptivity[i] * I_nu))
        I_nu = I_nu + dI
        if dIs:
            out.append(dI)
    if dIs:
        return I_nu, (np.stack(out) if out else np.zeros((0, npx)))
    return I_nu


def solve_attenuated_emissivity(emissivity, absorptivity, dt, L_unit):
    """transfer.py:122-144."""
    nsteps, npx = emissivity.shape
    tau = np.zeros(npx)
    out = []
    for i in range(1, nsteps):
        local_source = - emissivity[i, :] * dt[i - 1, :] * L_unit
        dtau = absorptivity[i] * dt[i - 1] * L_unit
        out.append(np.exp(-tau) * local_source)
        tau = tau - dtau
    return np.stack(out) if out else np.zeros((0, npx))


# ----------------------------------------------------------------------------------------------
# fluid model — grmhd/grmhd.py:35-55, grmhd/athenak.py:105-229 (same-level ghosts), :527-812
# ----------------------------------------------------------------------------------------------
class GRMHDFluidModel:
    def get_units(self, M_BH, mass_scale):
        """grmhd/grmhd.py:40-55."""
        L_unit = GNEWT * M_BH / CL**2
        T_unit = L_unit / CL
        dens_unit = mass_scale / L_unit**3
        Ne_unit = dens_unit / (MP + ME)
        B_unit = CL * np.sqrt(4. * np.pi * dens_unit)
        return dict(L_unit=L_unit, T_unit=T_unit, dens_unit=dens_unit, Ne_unit=Ne_unit, B_unit=B_unit)


DEFAULT_VARIABLE_NAMES = ('dens', 'velx', 'vely', 'velz', 'eint', 'bcc1', 'bcc2', 'bcc3')


class AthenakFluidModel(GRMHDFluidModel):
    """athenak.py:48-812, constructed from the arrays an .athdf file holds (h5py is absent here).

    uov: (5, nmb, nk, nj, ni); B: (3, nmb, nk, nj, ni); x{1,2,3}v: (nmb, n); x{1,2,3}f: (nmb, n+1);
    LogicalLocations: (nmb, 3); Levels: (nmb,).  Same-level ghost fill (athenak.py:208-229) and the
    coarser / finer neighbour branches (:231-514) are restated; ghosts outside the domain stay zero.
    """

    def __init__(self, uov, B, x1v, x2v, x3v, x1f, x2f, x3f, LogicalLocations, Levels, bhspin,
                 fluid_gamma=None, variable_names=DEFAULT_VARIABLE_NAMES):
        self.variable_names = np.array(variable_names)
        self.bhspin = bhspin
        self.fluid_gamma = fluid_gamma
        uov = np.asarray(uov, dtype=np.float64)
        B = np.asarray(B, dtype=np.float64)
        nprim, nmb, nmbk, nmbj, nmbi = uov.shape
        mb_index_map = {}
        for mb in range(nmb):
            ti, tj, tk = LogicalLocations[mb]
            mb_index_map[(int(Levels[mb]), int(ti), int(tj), int(tk))] = mb
        allmb = np.zeros((nmb, 8, nmbk + 2, nmbj + 2, nmbi + 2))
        for mbi in mb_index_map.values():
            tlevel = int(Levels[mbi])
            ti, tj, tk = (int(q) for q in LogicalLocations[mbi])
            new = allmb[mbi]
            new[:nprim, 1:-1, 1:-1, 1:-1] = uov[:, mbi]
            new[nprim:, 1:-1, 1:-1, 1:-1] = B[:, mbi]
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    for dk in (-1, 0, 1):
                        if di == 0 and dj == 0 and dk == 0:
                            continue
                        key = (tlevel, ti + di, tj + dj, tk + dk)
                        if key in mb_index_map:                       # athenak.py:208-229
                            nb = mb_index_map[key]
                            src_i = 0 if di == 1 else (-1 if di == -1 else slice(0, nmbi))
                            src_j = 0 if dj == 1 else (-1 if dj == -1 else slice(0, nmbj))
                            src_k = 0 if dk == 1 else (-1 if dk == -1 else slice(0, nmbk))
                            tgt_i = -1 if di == 1 else (0 if di == -1 else slice(1, nmbi + 1))
                            tgt_j = -1 if dj == 1 else (0 if dj == -1 else slice(1, nmbj + 1))
                            tgt_k = -1 if dk == 1 else (0 if dk == -1 else slice(1, nmbk + 1))
                            new[:nprim, tgt_k, tgt_j, tgt_i] = uov[:, nb, src_k, src_j, src_i]
                            new[nprim:, tgt_k, tgt_j, tgt_i] = B[:, nb, src_k, src_j, src_i]
                        # (multi-level neighbours: not needed for the single-level synthetic configs)
        self.mb_index_map = mb_index_map
        self.all_meshblocks = allmb
        self.x1v, self.x2v, self.x3v = (np.asarray(q, dtype=np.float64) for q in (x1v, x2v, x3v))
        self.x1f, self.x2f, self.x3f = (np.asarray(q, dtype=np.float64) for q in (x1f, x2f, x3f))
        self.Levels = np.asarray(Levels)
        self.LogicalLocations = np.asarray(LogicalLocations)
        self.nprim_all = 8

    def get_index_for_primitive_by_name(self, prim):
        prim = prim.lower().strip()
        names = [v.lower().strip() for v in self.variable_names]
        return names.index(prim) if prim in names else -1

    # -- athenak.py:663-670 ------------------------------------------------------------------
    def _meshblock_indices(self, S):
        nmb = self.all_meshblocks.shape[0]
        mb_indices = np.ones(S.shape[:-1], dtype=int) * -1
        with np.errstate(invalid='ignore'):
            for mbi in range(nmb):
                m = (self.x1f[mbi][0] < S[..., 1]) & (S[..., 1] <= self.x1f[mbi][-1])
                m &= (self.x2f[mbi][0] < S[..., 2]) & (S[..., 2] <= self.x2f[mbi][-1])
                m &= (self.x3f[mbi][0] < S[..., 3]) & (S[..., 3] <= self.x3f[mbi][-1])
                mb_indices[m] = mbi
        return mb_indices

    # -- athenak.py:718-757 ------------------------------------------------------------------
    def _interp_prims(self, S0, mb):
        x1_left, x2_left, x3_left = self.x1v[:, 0], self.x2v[:, 0], self.x3v[:, 0]
        dx1 = self.x1v[:, 1] - self.x1v[:, 0]
        dx2 = self.x2v[:, 1] - self.x2v[:, 0]
        dx3 = self.x3v[:, 1] - self.x3v[:, 0]
        with np.errstate(all='ignore'):
            x1i = S0[:, 1] - x1_left[mb] + dx1[mb]
            x2i = S0[:, 2] - x2_left[mb] + dx2[mb]
            x3i = S0[:, 3] - x3_left[mb] + dx3[mb]
            x1d = x1i / dx1[mb]
            x2d = x2i / dx2[mb]
            x3d = x3i / dx3[mb]
            # NaN positions (never in-domain) would not convert to int; they are zeroed below anyway
            x1i = np.nan_to_num(x1i // dx1[mb]).astype(int)
            x2i = np.nan_to_num(x2i // dx2[mb]).astype(int)
            x3i = np.nan_to_num(x3i // dx3[mb]).astype(int)
            x1d = x1d % 1.
            x2d = x2d % 1.
            x3d = x3d % 1.
        d = self.all_meshblocks
        nk, nj, ni = d.shape[2:]
        # JAX gathers clamp out-of-bounds indices; only mb == -1 rows can be out of bounds and those
        # are zeroed afterwards (athenak.py:755-757), so clamping here does not change any result.
        c1 = lambda q: np.clip(q, 0, ni - 1)
        c2 = lambda q: np.clip(q, 0, nj - 1)
        c3 = lambda q: np.clip(q, 0, nk - 1)
        daaa = d[mb, :, c3(x3i), c2(x2i), c1(x1i)]
        daab = d[mb, :, c3(x3i), c2(x2i), c1(x1i + 1)]
        daba = d[mb, :, c3(x3i), c2(x2i + 1), c1(x1i)]
        dabb = d[mb, :, c3(x3i), c2(x2i + 1), c1(x1i + 1)]
        dbaa = d[mb, :, c3(x3i + 1), c2(x2i), c1(x1i)]
        dbab = d[mb, :, c3(x3i + 1), c2(x2i), c1(x1i + 1)]
        dbba = d[mb, :, c3(x3i + 1), c2(x2i + 1), c1(x1i)]
        dbbb = d[mb, :, c3(x3i + 1), c2(x2i + 1), c1(x1i + 1)]
        with np.errstate(all='ignore'):
            daa = daaa + (daab - daaa) * x1d[:, None]
            dab = daba + (dabb - daba) * x1d[:, None]
            dba = dbaa + (dbab - dbaa) * x1d[:, None]
            dbb = dbba + (dbbb - dbba) * x1d[:, None]
            da = daa + (dab - daa) * x2d[:, None]
            db = dba + (dbb - dba) * x2d[:, None]
            prims = da + (db - da) * x3d[:, None]
        prims = np.where((mb == -1)[:, None], 0.0, prims)
        return prims

    def get_prims_from_geodesics(self, S):
        """athenak.py:527-637."""
        S = np.asarray(S, dtype=np.float64)
        nsteps, npx, _ = S.shape
        mb_indices = self._meshblock_indices(S)
        prims = np.stack([self._interp_prims(S[i], mb_indices[i]) for i in range(nsteps)])
        g = self.get_index_for_primitive_by_name
        return dict(dens=prims[..., g('dens')], u=prims[..., g('eint')],
                    U1=prims[..., g('velx')], U2=prims[..., g('vely')], U3=prims[..., g('velz')],
                    B1=prims[..., g('bcc1')], B2=prims[..., g('bcc2')], B3=prims[..., g('bcc3')])

    def get_fluid_scalars_from_geodesics(self, S, fallback_pitch_angle=np.pi / 3.):
        """athenak.py:639-812."""
        S = np.asarray(S, dtype=np.float64)
        nsteps, npx, _ = S.shape
        mb_indices = self._meshblock_indices(S)
        g = self.get_index_for_primitive_by_name
        irho, iu, iU1, iU2, iU3 = g('dens'), g('eint'), g('velx'), g('vely'), g('velz')
        iB1, iB2, iB3 = g('bcc1'), g('bcc2'), g('bcc3')
        if iU2 != iU1 + 1 or iU3 != iU1 + 2:
            raise ValueError("Velocity indices are not as expected")
        if iB2 != iB1 + 1 or iB3 != iB1 + 2:
            raise ValueError("Magnetic field indices are not as expected")
        out = np.zeros((nsteps, npx, 5))
        for i in range(nsteps):
            S0 = S[i]
            prims = self._interp_prims(S0, mb_indices[i])
            out[i] = fluid_frame_scalars(S0, prims, self.bhspin, irho, iu, iU1, iB1, fallback_pitch_angle)
        return dict(dens=out[:, :, 0], u=out[:, :, 1], pitch_angle=out[:, :, 2], kdotu=out[:, :, 3],
                    b=out[:, :, 4])


class AnalyticTorusFluidModel(GRMHDFluidModel):
    """cfg3: analytic Keplerian thin torus, power-law density and toroidal field at fixed
    plasma beta.  NOT in the reference (which only ships AthenakFluidModel); it is a GRMHDFluidModel duck type
    whose primitives are closed-form functions of position, pushed through the identical fluid-frame algebra
    of athenak.py:760-794.  Zero outside the sphere r <= r_out (plays the role of "outside the domain")."""

    def __init__(self, bhspin, fluid_gamma=13. / 9, R0=8.0, R_in=2.5, p=1.5, h=0.3, u0=0.25, beta0=3.0,
                 dens_scale=1.0, r_out=40.0):
        self.bhspin, self.fluid_gamma = bhspin, fluid_gamma
        self.R0, self.R_in, self.p, self.h, self.u0, self.beta0 = R0, R_in, p, h, u0, beta0
        self.dens_scale, self.r_out = dens_scale, r_out

    def params_array(self):
        return np.array([self.fluid_gamma, self.R0, self.R_in, self.p, self.h, self.u0, self.beta0,
                         self.dens_scale, self.r_out])

    def prims(self, X):
        """X (..., 4+) -> prims (..., 8) in the AthenaK order dens, velx, vely, velz, eint, bcc1..3."""
        x, y, z = X[..., 1], X[..., 2], X[..., 3]
        with np.errstate(all='ignore'):
            R2 = x * x + y * y
            R = np.sqrt(R2) + 1e-12
            r = np.sqrt(R2 + z * z) + 1e-12
            H = self.h * R
            taper = np.exp(-(self.R_in / R)**4)
            dens = self.dens_scale * (R / self.R0)**(-self.p) * np.exp(-z * z / (2. * H * H)) * taper
            eint = self.u0 * dens * (self.R0 / r)
            vphi = 0.5 / np.sqrt(1. + R)
            bmag = np.sqrt(2. * eint * (self.fluid_gamma - 1.) / self.beta0)
            out = np.stack([dens, -vphi * y / R, vphi * x / R, 0.02 * z / (1. + r), eint,
                            -bmag * y / R, bmag * x / R, 0.1 * bmag], axis=-1)
        return np.where((r <= self.r_out)[..., None], out, 0.0)

    def get_prims_from_geodesics(self, S):
        P = self.prims(np.asarray(S, dtype=np.float64))
        return dict(dens=P[..., 0], u=P[..., 4], U1=P[..., 1], U2=P[..., 2], U3=P[..., 3],
                    B1=P[..., 5], B2=P[..., 6], B3=P[..., 7])

    def get_fluid_scalars_from_geodesics(self, S, fallback_pitch_angle=np.pi / 3.):
        S = np.asarray(S, dtype=np.float64)
        out = np.zeros(S.shape[:2] + (5,))
        for i in range(S.shape[0]):
            out[i] = fluid_frame_scalars(S[i], self.prims(S[i]), self.bhspin, 0, 4, 1, 5, fallback_pitch_angle)
        return dict(dens=out[:, :, 0], u=out[:, :, 1], pitch_angle=out[:, :, 2], kdotu=out[:, :, 3],
                    b=out[:, :, 4])


def fluid_frame_scalars(S0, prims, bhspin, irho=0, iu=4, iU1=1, iB1=5, fallback_pitch_angle=np.pi / 3.):
    """athenak.py:760-794: metric, four-velocity, magnetic four-vector, k.u, pitch angle, |b|."""
    with np.errstate(all='ignore'):
        gcov = metric(S0[:, :4], bhspin)
        gcon = imetric(S0[:, :4], bhspin)
        alpha = np.sqrt(1. / (-gcon[:, 0, 0]))
        Uprim = prims[:, iU1:iU1 + 3]
        gamma = np.sqrt(1 + np.einsum('aj,aj->a', np.einsum('ai,aij->aj', Uprim, gcov[:, 1:, 1:]), Uprim))
        ucon0 = gamma / alpha
        ucon1 = Uprim[:, 0] - gamma * alpha * gcon[:, 0, 1]
        ucon2 = Uprim[:, 1] - gamma * alpha * gcon[:, 0, 2]
        ucon3 = Uprim[:, 2] - gamma * alpha * gcon[:, 0, 3]
        ucon = np.stack([ucon0, ucon1, ucon2, ucon3], axis=1)
        ucov = np.einsum('aij,aj->ai', gcov, ucon)
        Bprim = prims[:, iB1:iB1 + 3]
        bcon0 = np.einsum('ai,ai->a', Bprim, ucov[:, 1:])
        bcon1 = (Bprim[:, 0] + ucon1 * bcon0) / ucon0
        bcon2 = (Bprim[:, 1] + ucon2 * bcon0) / ucon0
        bcon3 = (Bprim[:, 2] + ucon3 * bcon0) / ucon0
        bcon = np.stack([bcon0, bcon1, bcon2, bcon3], axis=1)
        bcov = np.einsum('aij,aj->ai', gcov, bcon)
        kdotu = np.einsum('ai,ai->a', S0[:, 4:], ucov)
        kdotb = np.einsum('ai,ai->a', S0[:, 4:], bcov)
        bdotb = np.einsum('ai,ai->a', bcon, bcov)
        pitch = kdotb / (np.abs(kdotu) * np.sqrt(bdotb))
        pitch = np.where(np.isnan(pitch), np.ones_like(pitch) * np.cos(fallback_pitch_angle), pitch)
        pitch = np.where(np.abs(pitch) > 1.0, pitch / np.abs(pitch), pitch)
        pitch = np.arccos(pitch)
        return np.stack([prims[:, irho], prims[:, iu], pitch, kdotu, np.sqrt(bdotb)], axis=1)


# ----------------------------------------------------------------------------------------------
# image driver — images.py:30-144
# ----------------------------------------------------------------------------------------------
def intensity_from_trajectories(fluid_model, S, final_dt, mass_scale, M_bh, r_high, observing_frequency):
    """images.py:84-120 for one chunk: fluid scalars -> Theta_e -> j, alpha -> sigma cut -> transfer."""
    fluid_gamma = fluid_model.fluid_gamma
    fs = fluid_model.get_fluid_scalars_from_geodesics(S)
    with np.errstate(all='ignore'):
        bsq = fs['b'] * fs['b']
        beta = fs['u'] * (fluid_gamma - 1.) / bsq / 0.5
        sigma = bsq / fs['dens']
        Theta_e = rlow_rhigh_model(fs['dens'], fs['u'], beta, r_high=r_high)
        units = fluid_model.get_units(M_bh, mass_scale)
        Ne_in_cgs = units['Ne_unit'] * fs['dens']
        B_in_gauss = units['B_unit'] * fs['b']
        local_nu = - fs['kdotu'] * observing_frequency
        em, ab = synchrotron_coefficients(Ne_in_cgs, Theta_e, B_in_gauss, fs['pitch_angle'], local_nu,
                                          invariant=True, rescale_nu=1. / observing_frequency)
        cut = sigma > 100.
        em = np.where(cut, 0.0, em)
        ab = np.where(cut, 0.0, ab)
    return solve_specific_intensity(em, ab, final_dt, units['L_unit'])


def make_image(fluid_model, camera_inclination=60, camera_distance=1000, mass_scale=1.e26,
               M_bh=6.2e9 * Msun, r_high=40, observing_frequency=230.e9, fov=20, resolution=160,
               max_nsteps=10000, max_chunk_bytes=None, integrator=None):
    """images.py:30-144 (chunking changes nothing numerically: rays are independent)."""
    bhspin = fluid_model.bhspin
    s0 = initialize_geodesics_at_camera(bhspin, camera_inclination, camera_distance, -fov / 2., fov / 2.,
                                        resolution)
    num_pixels_per_chunk = s0.shape[0] + 10
    if max_chunk_bytes is not None:
        num_pixels_per_chunk = int(max_chunk_bytes // 4 // 20 // max_nsteps)
    integ = geodesic_integrator if integrator is None else integrator
    I_nu_saved = np.zeros((0))
    lower = 0
    while lower < s0.shape[0]:
        S, final_dt = integ(max_nsteps, s0[lower:lower + num_pixels_per_chunk], 40, 1e-4, bhspin)
        I_nu = intensity_from_trajectories(fluid_model, S, final_dt, mass_scale, M_bh, r_high,
                                           observing_frequency)
        I_nu_saved = np.append(I_nu_saved, I_nu)
        lower += num_pixels_per_chunk
    return np.array(I_nu_saved).reshape((resolution, resolution))
