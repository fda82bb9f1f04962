// _quadratic / _Nullify of /root/reference/mahakala/geodesics.py:58-85 for an arbitrary covariant metric
// (NVRTC-safe: shared by the built-in camera kernels and run-time compiled metric plugins).
#pragma once

namespace mk {

// nullify (geodesics.py:73-83) given the covariant metric at x: rescale the spatial part of v so that
// g_mn v^m v^n = 0 keeping v^t; writes the 8-vector state.
__device__ __forceinline__ void nullify_with_metric(const double gm[4][4], const double x[4], const double v[4], double s[8])
{
    double A = v[0] * gm[0][0] * v[0];
    double b = (v[1] * gm[1][0] + v[2] * gm[2][0] + v[3] * gm[3][0]) * v[0];
    double C = 0.0;
#pragma unroll
    for (int j = 1; j < 4; j++) C += (v[1] * gm[1][j] + v[2] * gm[2][j] + v[3] * gm[3][j]) * v[j];
    // _quadratic
    double bb = b * b, AC = A * C;
    bool close = fabs(bb - AC) <= (1e-8 + 1e-5 * fabs(AC));        // jnp.isclose defaults
    double dd = close ? 0.0 : bb - AC;
    double bs = (b < 0.0) ? 0.0 : ((b != b) ? b : 1.0);            // heaviside(b, 1)
    double D = -(b + bs * sqrt(dd));
    double x1 = D / A, x2 = C / D;
    double d1 = fmin(x1, x2), d2 = fmax(x1, x2);
    if (x1 != x1 || x2 != x2) d1 = d2 = nan("");                   // jnp.minimum/maximum propagate NaN
    double S = (d1 > 0.0) ? d1 : ((d2 > 0.0) ? d2 : nan(""));
    s[0] = x[0]; s[1] = x[1]; s[2] = x[2]; s[3] = x[3];
    s[4] = v[0]; s[5] = v[1] / S; s[6] = v[2] / S; s[7] = v[3] / S;
}

}  // namespace mk
