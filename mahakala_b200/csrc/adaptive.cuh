// OPTIONAL integrator: embedded Dormand-Prince 5(4) with step-size control ("adaptive / embedded
// RK as an option").  Not in the reference: /root/reference/mahakala/geodesics.py:246-269 takes fixed-rule steps
// dt = -(r - r_H)/div with classical RK4 (:317-336).  That rule is specific to Kerr (it needs r and r_H to size the
// step) and spends ~500-3800 steps per ray at div = 40; a user-registered spacetime has no reason to be well resolved by
// it.  Here the step follows the local truncation error instead.  Kept deliberately close to the reference's loop:
//   * same termination test: a ray is alive while tol <= radius - r_H <= far (1500 in the reference, geodesics.py:250-252);
//   * same freeze semantics: a step whose end point fails the test is rejected and the ray stays where it was;
//   * same direction: the affine parameter runs backwards (h < 0);
//   * |h| <= cap (r - r_H): captured rays still approach the horizon geometrically (cap = 1/2: ~log2 steps), never jump
//     across it.
// 6 acceleration evaluations per accepted step (the 7th stage is the first of the next step, FSAL).
// NVRTC-safe; host-compilable for tests/host_harness.
#pragma once
#include "integrate.cuh"

namespace mk {

struct AdaptiveRule {
    double rtol, atol;      // per-step error bound: |err(x)| <= atol + rtol |x|, |err(k)| <= atol + rtol |k| (Euclidean)
    double tol, far, rH;    // alive while tol <= radius - rH <= far
    double cap;             // |h| <= cap (radius - rH)
    double div0;            // first trial step (radius - rH) / div0 (the reference's rule as the starting guess)

    MK_HD bool alive(double r) const
    {
        const double m = r - rH;
        return (m >= tol) & (m <= far);         // NaN fails both
    }
};

// One trial step of size h from y (k1 = f(y)): 5th-order solution yn, its derivative k7 = f(yn), and the scaled error
// estimate err (<= 1: accept).
template <class Metric>
MK_HD void dopri5_trial(const Metric& g, const AdaptiveRule& R, const double y[8], const double k1[8], double h,
                        double yn[8], double k7[8], double& err)
{
    double k2[8], k3[8], k4[8], k5[8], k6[8], ys[8];
    auto f = [&](const double* s, double* k) {
#pragma unroll
        for (int i = 0; i < 4; i++) k[i] = s[4 + i];
        g.accel(s, s + 4, k + 4);
    };
#pragma unroll
    for (int i = 0; i < 8; i++) ys[i] = fma(h * (1.0 / 5.0), k1[i], y[i]);
    f(ys, k2);
#pragma unroll
    for (int i = 0; i < 8; i++) ys[i] = fma(h, fma(3.0 / 40.0, k1[i], (9.0 / 40.0) * k2[i]), y[i]);
    f(ys, k3);
#pragma unroll
    for (int i = 0; i < 8; i++)
        ys[i] = fma(h, fma(44.0 / 45.0, k1[i], fma(-56.0 / 15.0, k2[i], (32.0 / 9.0) * k3[i])), y[i]);
    f(ys, k4);
#pragma unroll
    for (int i = 0; i < 8; i++)
        ys[i] = fma(h, fma(19372.0 / 6561.0, k1[i], fma(-25360.0 / 2187.0, k2[i], fma(64448.0 / 6561.0, k3[i],
                   (-212.0 / 729.0) * k4[i]))), y[i]);
    f(ys, k5);
#pragma unroll
    for (int i = 0; i < 8; i++)
        ys[i] = fma(h, fma(9017.0 / 3168.0, k1[i], fma(-355.0 / 33.0, k2[i], fma(46732.0 / 5247.0, k3[i],
                   fma(49.0 / 176.0, k4[i], (-5103.0 / 18656.0) * k5[i])))), y[i]);
    f(ys, k6);
#pragma unroll
    for (int i = 0; i < 8; i++)
        yn[i] = fma(h, fma(35.0 / 384.0, k1[i], fma(500.0 / 1113.0, k3[i], fma(125.0 / 192.0, k4[i],
                   fma(-2187.0 / 6784.0, k5[i], (11.0 / 84.0) * k6[i])))), y[i]);
    f(yn, k7);
    // error of the position 3-vector against the length of the position, error of the wavevector against its length:
    // component-wise relative bounds would force tiny steps whenever a single coordinate crosses zero
    double e2x = 0.0, e2k = 0.0, x2 = 0.0, kk2 = 0.0;
#pragma unroll
    for (int i = 1; i < 8; i++) {           // t (i = 0) feeds nothing back: not controlled
        const double e = h * fma(71.0 / 57600.0, k1[i], fma(-71.0 / 16695.0, k3[i], fma(71.0 / 1920.0, k4[i],
                             fma(-17253.0 / 339200.0, k5[i], fma(22.0 / 525.0, k6[i], (-1.0 / 40.0) * k7[i])))));
        const double m = fmax(fabs(y[i]), fabs(yn[i]));
        if (i < 4) { e2x = fma(e, e, e2x); x2 = fma(m, m, x2); }
        else { e2k = fma(e, e, e2k); kk2 = fma(m, m, kk2); }
    }
    const double qx = sqrt(e2x) / fma(R.rtol, sqrt(x2), R.atol);
    const double qk = sqrt(e2k) / fma(R.rtol, sqrt(kk2), R.atol);
    err = (qx > qk || qx != qx) ? qx : qk;                  // a NaN poisons the estimate -> rejected
    if (qk != qk) err = qk;
}

// One ray from y to its end.  Returns the radius of the final state (the classifier: ~r_H + tol for a captured ray,
// hundreds of M for an escaped one); y is overwritten with the final state.
template <class Metric>
MK_HD double integrate_one_adaptive(const Metric& g, const AdaptiveRule& R, double (&y)[8], int N, int& nsteps, int& nrejected)
{
    nsteps = 0;
    nrejected = 0;
    double r = g.radius(y);
    if (!R.alive(r) || N <= 0) return r;
    double k1[8];
#pragma unroll
    for (int i = 0; i < 4; i++) k1[i] = y[4 + i];
    g.accel(y, y + 4, k1 + 4);
    // The controlled quantity is eta = |h| / (r - r_H), the step in units of the distance to the horizon, not h itself:
    // towards the horizon the admissible step shrinks geometrically with r - r_H, and a controller that carries h over
    // from one step to the next overshoots after every accepted step (measured: one rejection per accepted step there).
    double eta = 1.0 / R.div0;
    for (;;) {
        const double h = -fmin(eta, R.cap) * (r - R.rH);
        double yn[8], k7[8], err;
        dopri5_trial(g, R, y, k1, h, yn, k7, err);
        if (!(err <= 1.0)) {
            // too large (or not a number): shrink and retry from the same point
            nrejected++;
            eta = fmin(eta, R.cap) * ((err == err) ? fmax(0.2, 0.9 * pow(err, -0.2)) : 0.2);
            if (!(eta > 1e-12) || nrejected > 64 + 8 * nsteps) break;      // cannot proceed: frozen
            continue;
        }
        const double rn = g.radius(yn);
        if (!R.alive(rn)) break;                // end point outside the live range: rejected, ray frozen at y
#pragma unroll
        for (int i = 0; i < 8; i++) { y[i] = yn[i]; k1[i] = k7[i]; }
        r = rn;
        nsteps++;
        if (nsteps == N) break;
        eta = fmin(eta, R.cap) * ((err > 1e-10) ? fmin(5.0, 0.9 * pow(err, -0.2)) : 5.0);
    }
    return r;
}

struct AdaptiveArgs {
    const double* s0;       // (npx, 8)
    long npx;
    int N;                  // cap on accepted steps
    AdaptiveRule rule;
    double* final_state;    // (npx, 8) or null
    int* nsteps;            // (npx,) accepted steps, or null
    int* nrejected;         // (npx,) rejected trial steps, or null
    double* r_last;         // (npx,) radius of the final state, or null
};

// one ray per thread (rays are short -- tens of steps -- so there is no queue / refill machinery here)
template <class Metric>
__device__ __forceinline__ void integrate_adaptive_body(const Metric& g, const AdaptiveArgs& A)
{
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < A.npx; p += (long)gridDim.x * blockDim.x) {
        double y[8];
        const double2* src = reinterpret_cast<const double2*>(A.s0 + p * 8);
#pragma unroll
        for (int i = 0; i < 4; i++) { double2 v = src[i]; y[2 * i] = v.x; y[2 * i + 1] = v.y; }
        int ns, nr;
        const double r = integrate_one_adaptive(g, A.rule, y, A.N, ns, nr);
        if (A.final_state) {
            double2* dst = reinterpret_cast<double2*>(A.final_state + p * 8);
#pragma unroll
            for (int i = 0; i < 4; i++) dst[i] = make_double2(y[2 * i], y[2 * i + 1]);
        }
        if (A.nsteps) A.nsteps[p] = ns;
        if (A.nrejected) A.nrejected[p] = nr;
        if (A.r_last) A.r_last[p] = r;
    }
}

}  // namespace mk
