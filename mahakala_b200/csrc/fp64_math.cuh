// FP64 building blocks for the ray kernels (sm_100a).
//
// The integrator is bound by the FP64 FMA pipe (64 DFMA/clk/SM on B200), so divisions and square
// roots are expressed as one MUFU seed (RCP64H / RSQ64H, ~2^-22 relative) plus FMA-only Newton steps
// that run on the same pipe with no branches or slow paths.  Accuracy is <= ~2 ulp (checked against IEEE
// results by mk_fast_math_probe / tests/test_geodesics_gpu.py::test_fast_math_accuracy),
// far inside the 1e-9 trajectory tolerance; inputs on this path are normal, finite and positive
// (radii, metric denominators), so the IEEE special-case handling of '/' and sqrt() is not needed.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

namespace mk {

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ double rcp_seed(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

__device__ __forceinline__ double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// 1/x to ~1 ulp: MUFU seed (relative error e0 <= ~2^-20) + one cubically convergent step
//   y1 = y0 (1 + e + e^2),  e = 1 - x y0      ->  error e0^3 <= 2^-60          (3 DFMA)
__device__ __forceinline__ double fast_rcp(double x)
{
    double y = rcp_seed(x);
    double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// sqrt(x) and 1/sqrt(x) together: MUFU seed + one cubically convergent step for the reciprocal root
//   y1 = y0 (1 + e/2 + 3 e^2/8),  e = 1 - x y0^2                                  (5 FP64 ops)
// then s = x y1 with one residual correction s += (x - s^2) y1/2                   (4 FP64 ops)
__device__ __forceinline__ void fast_sqrt_rsqrt(double x, double& s, double& rs)
{
    double y = rsqrt_seed(x);
    double e = fma(-(x * y), y, 1.0);
    double p = fma(0.375, e, 0.5) * e;
    y = fma(y, p, y);
    double g = x * y;
    double d = fma(-g, g, x);
    s = fma(d, 0.5 * y, g);
    rs = y;
}

__device__ __forceinline__ double fast_sqrt(double x)
{
    double s, rs;
    fast_sqrt_rsqrt(x, s, rs);
    return s;
}

// The same without the final residual correction: sqrt(x) = x * rsqrt(x) to <= ~3 ulp in 6 FP64 ops.  Used
// inside the geodesic right-hand side, whose inputs already carry the rounding of the previous stage.
__device__ __forceinline__ void quick_sqrt_rsqrt(double x, double& s, double& rs)
{
    double y = rsqrt_seed(x);
    double e = fma(-(x * y), y, 1.0);
    double p = fma(0.375, e, 0.5) * e;
    y = fma(y, p, y);
    s = x * y;
    rs = y;
}

__device__ __forceinline__ double quick_sqrt(double x)
{
    double s, rs;
    quick_sqrt_rsqrt(x, s, rs);
    return s;
}

// a / b with b's reciprocal refined and one residual correction (≈ correctly rounded).
__device__ __forceinline__ double fast_div(double a, double b)
{
    double y = fast_rcp(b);
    double q = a * y;
    double e = fma(-q, b, a);
    return fma(e, y, q);
}

}  // namespace mk
