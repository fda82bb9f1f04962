// FP64 building blocks for the ray kernels (sm_100a).
//
// The integrator is bound by the FP64 FMA pipe (64 DFMA/clk/SM on B200), so divisions and square
// roots are expressed as one MUFU seed (RCP64H / RSQ64H, ~2^-22 relative) plus FMA-only Newton steps
// that run on the same pipe with no branches or slow paths.  Accuracy after two steps is <= ~1 ulp,
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

// 1/x to ~1 ulp: seed + two Newton steps (4 DFMA).
__device__ __forceinline__ double fast_rcp(double x)
{
    double y = rcp_seed(x);
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// sqrt(x) and 1/sqrt(x) together (Goldschmidt-style coupled iteration + one residual correction).
__device__ __forceinline__ void fast_sqrt_rsqrt(double x, double& s, double& rs)
{
    double y = rsqrt_seed(x);
    double g = x * y;         // ~ sqrt(x)
    double h = 0.5 * y;       // ~ 1 / (2 sqrt(x))
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    double d = fma(-g, g, x);
    s = fma(d, h, g);
    rs = h + h;
}

__device__ __forceinline__ double fast_sqrt(double x)
{
    double s, rs;
    fast_sqrt_rsqrt(x, s, rs);
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
