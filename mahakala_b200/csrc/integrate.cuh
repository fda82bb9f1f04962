// Per-ray stepping shared by the integrate-only and the fused render kernels.
//
// Semantics restated from /root/reference/mahakala/geodesics.py:
//   step rule      :249-252 / :258-261   dt = -(radius_cal - r_H)/div ; 0 if NaN, |dt|*div < tol, |dt|*div > 1500
//   RK4            :317-336
//   reject/freeze  :264-267             if rule(candidate) == 0 the step is rejected and the ray is frozen
// A frozen ray re-proposes the same rejected step forever in the reference's lax.scan; here it retires.
#pragma once
#include "fp64_math.cuh"

namespace mk {

struct StepRule {
    double div, inv_div, tol, rH;

    // returns dt for a state whose step-rule radius is r
    MK_HD double operator()(double r) const
    {
        double num = rH - r;                       // -(r - r_H)
        double q = num * inv_div;                  // division by div, residual-corrected
        q = fma(fma(-q, div, num), inv_div, q);
        double m = fabs(q) * div;
        // geodesics.py:250-252: zero if NaN, |dt| div < tol or |dt| div > 1500.  A NaN fails both ordered
        // comparisons below, so the negated conjunction covers all three cases with two compares.
        bool keep = (m >= tol) & (m <= 1500.0);
        return keep ? q : 0.0;
    }
};

// One classical RK4 step of the 8-vector (x^m, v^m) with the plugin's acceleration (geodesics.py:317-336).
//
// The system is second order (dx/dl = v, dv/dl = a(x, v)), so the position half of the RK4 combination collapses:
// with v2 = v + h/2 a1, v3 = v + h/2 a2, v4 = v + h a3
//     x' = x + h/6 (v + 2 v2 + 2 v3 + v4) = x + h v + h^2/6 (a1 + a2 + a3)
//     v' = v + h/6 (a1 + 2 a2 + 2 a3 + a4)
// -- the same numbers as the literal form up to rounding, with 52 instead of 56 FP64 operations around the four
// acceleration calls.  Split in two so that the fused render kernel can sample the snapshot between the first
// stage (whose metric functions f, l it reuses) and the rest of the step.
struct Rk4Stage1 {
    double a1[4];
};

template <class Metric>
MK_HD void rk4_rest(const Metric& g, const double s[8], const double a1[4], double dt, double out[8])
{
    const double hdt = 0.5 * dt;
    double x2[4], v2[4], a2[4], a3[4], a4[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        x2[i] = fma(hdt, s[4 + i], s[i]);
        v2[i] = fma(hdt, a1[i], s[4 + i]);
    }
    g.accel(x2, v2, a2);                                   // stage 2
    double x3[4], v3[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        x3[i] = fma(hdt, v2[i], s[i]);
        v3[i] = fma(hdt, a2[i], s[4 + i]);
    }
    g.accel(x3, v3, a3);                                   // stage 3
    double x4[4], v4[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        x4[i] = fma(dt, v3[i], s[i]);
        v4[i] = fma(dt, a3[i], s[4 + i]);
    }
    g.accel(x4, v4, a4);                                   // stage 4
    const double sdt = dt * (1.0 / 6.0);
    const double sdt2 = dt * sdt;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double a23 = a2[i] + a3[i];
        double a123 = a1[i] + a23;
        double xn = fma(sdt2, a123, fma(dt, s[4 + i], s[i]));
        double vn = fma(sdt, (a123 + a23) + a4[i], s[4 + i]);
        out[i] = xn;
        out[4 + i] = vn;
    }
}

template <class Metric>
MK_HD void rk4_step(const Metric& g, const double s[8], double dt, double out[8],
                    const typename Metric::Cache* cache = nullptr)
{
    double a1[4];
    // stage 1 (the point cache of s comes from the step rule evaluated when s was accepted)
    g.accel(s, s + 4, a1, cache);
    rk4_rest(g, s, a1, dt, out);
}

}  // namespace mk
