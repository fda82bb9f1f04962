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
    __device__ __forceinline__ double operator()(double r) const
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

// One classical RK4 step of the 8-vector (x^m, v^m) with the plugin's acceleration.
template <class Metric>
__device__ __forceinline__ void rk4_step(const Metric& g, const double s[8], double dt, double out[8],
                                         const typename Metric::Cache* cache = nullptr)
{
    double acc[4], sum[8], tmp[8];
    const double hdt = 0.5 * dt;
    // stage 1 (the point cache of s comes from the step rule evaluated when s was accepted)
    g.accel(s, s + 4, acc, cache);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        sum[i] = s[4 + i];
        sum[4 + i] = acc[i];
        tmp[i] = fma(hdt, s[4 + i], s[i]);
        tmp[4 + i] = fma(hdt, acc[i], s[4 + i]);
    }
    // stage 2
    g.accel(tmp, tmp + 4, acc);
    {
        double t2[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            sum[i] = fma(2.0, tmp[4 + i], sum[i]);
            sum[4 + i] = fma(2.0, acc[i], sum[4 + i]);
            t2[i] = fma(hdt, tmp[4 + i], s[i]);
            t2[4 + i] = fma(hdt, acc[i], s[4 + i]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) tmp[i] = t2[i];
    }
    // stage 3
    g.accel(tmp, tmp + 4, acc);
    {
        double t2[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            sum[i] = fma(2.0, tmp[4 + i], sum[i]);
            sum[4 + i] = fma(2.0, acc[i], sum[4 + i]);
            t2[i] = fma(dt, tmp[4 + i], s[i]);
            t2[4 + i] = fma(dt, acc[i], s[4 + i]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) tmp[i] = t2[i];
    }
    // stage 4
    g.accel(tmp, tmp + 4, acc);
    const double sdt = dt * (1.0 / 6.0);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        out[i] = fma(sdt, sum[i] + tmp[4 + i], s[i]);
        out[4 + i] = fma(sdt, sum[4 + i] + acc[i], s[4 + i]);
    }
}

}  // namespace mk
