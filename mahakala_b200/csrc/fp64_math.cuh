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

// Every arithmetic helper of the ray kernels is __host__ __device__: the product only ever calls them on the device,
// but tests/host_harness compiles the SAME source for the host so that the CPU test suite (no GPU in the build
// container) can hold the kernels' arithmetic -- closed-form acceleration, RK4 step, step rule, emission chain -- to
// the CPU restatement of the reference.  On the host the MUFU seeds are replaced by library values truncated to 22 mantissa bits (the hardware's
// seeds are good to 2^-22..2^-23), so the Newton steps are exercised rather than bypassed.
#define MK_HD __host__ __device__ __forceinline__

#ifndef __CUDA_ARCH__
#ifndef __CUDACC_RTC__
#include <cmath>
#include <cstring>
#endif
#endif

namespace mk {

constexpr unsigned FULL_MASK = 0xffffffffu;

#ifndef __CUDA_ARCH__
static inline double host_truncate_seed(double y)
{
    unsigned long long b;
    std::memcpy(&b, &y, 8);
    b &= ~((1ULL << 30) - 1ULL);          // keep 22 mantissa bits
    std::memcpy(&y, &b, 8);
    return y;
}
#endif

MK_HD double rcp_seed(double x)
{
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return host_truncate_seed(1.0 / x);
#endif
}

MK_HD double rsqrt_seed(double x)
{
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return host_truncate_seed(1.0 / std::sqrt(x));
#endif
}

// bit-level helpers with host twins (the device intrinsics do not exist on the host)
MK_HD int hi_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
MK_HD int lo_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    long long b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL);
#endif
}
MK_HD int floor_to_int(double x)         // round towards minus infinity; NaN -> 0 (cvt.rmi.s32.f64)
{
#ifdef __CUDA_ARCH__
    return __double2int_rd(x);
#else
    return (x != x) ? 0 : (int)std::floor(x);
#endif
}
MK_HD double from_words(int hi, int lo)
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double x; std::memcpy(&x, &b, 8); return x;
#endif
}

// 1/x to ~1 ulp: MUFU seed (relative error e0 <= ~2^-20) + one cubically convergent step
//   y1 = y0 (1 + e + e^2),  e = 1 - x y0      ->  error e0^3 <= 2^-60          (3 DFMA)
MK_HD double fast_rcp(double x)
{
    double y = rcp_seed(x);
    double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// sqrt(x) and 1/sqrt(x) together: MUFU seed + one cubically convergent step for the reciprocal root
//   y1 = y0 (1 + e/2 + 3 e^2/8),  e = 1 - x y0^2                                  (5 FP64 ops)
// then s = x y1 with one residual correction s += (x - s^2) y1/2                   (4 FP64 ops)
MK_HD void fast_sqrt_rsqrt(double x, double& s, double& rs)
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

MK_HD double fast_sqrt(double x)
{
    double s, rs;
    fast_sqrt_rsqrt(x, s, rs);
    return s;
}

// The same without the final residual correction: sqrt(x) = x * rsqrt(x) to <= ~3 ulp in 6 FP64 ops.  Used
// inside the geodesic right-hand side, whose inputs already carry the rounding of the previous stage.
MK_HD void quick_sqrt_rsqrt(double x, double& s, double& rs)
{
    double y = rsqrt_seed(x);
    double e = fma(-(x * y), y, 1.0);
    double p = fma(0.375, e, 0.5) * e;
    y = fma(y, p, y);
    s = x * y;
    rs = y;
}

MK_HD double quick_sqrt(double x)
{
    double s, rs;
    quick_sqrt_rsqrt(x, s, rs);
    return s;
}

// Same accuracy class with the correction applied to both results independently:
//   g = x y0,  c = (1/2 + 3 e / 8) e,  sqrt = g + g c,  rsqrt = y0 + y0 c
// 6 FP64 operations for the pair, 5 for the square root alone, and the square root does not wait for the refined
// reciprocal root (one dependent operation less on the critical path of the geodesic right-hand side).
MK_HD void pair_sqrt_rsqrt(double x, double& s, double& rs)
{
    double y = rsqrt_seed(x);
    double g = x * y;
    double e = fma(-g, y, 1.0);
    double c = fma(0.375, e, 0.5) * e;
    s = fma(g, c, g);
    rs = fma(y, c, y);
}

MK_HD double sqrt_only(double x)
{
    double y = rsqrt_seed(x);
    double g = x * y;
    double e = fma(-g, y, 1.0);
    double c = fma(0.375, e, 0.5) * e;
    return fma(g, c, g);
}

// a / b with b's reciprocal refined and one residual correction (≈ correctly rounded).
MK_HD double fast_div(double a, double b)
{
    double y = fast_rcp(b);
    double q = a * y;
    double e = fma(-q, b, a);
    return fma(e, y, q);
}

// exp(-t) for t >= 0 (the synchrotron cut-off e^{-X^{1/3}}, transfer.py:66): k = rint(-t log2 e) by the 1.5 * 2^52
// trick, r = -t - k ln 2 in two FMAs (|r| <= 0.347), degree-13 Taylor polynomial (truncation 4e-18), 2^k by an
// integer add to the exponent field: 17 FP64 + 4 integer instructions against ~53 for exp().  <= ~1 ulp for
// t <= 707; beyond (result < 9e-308, where exp() would return subnormals) the result is flushed to 0.  NaN -> NaN.
MK_HD double fast_exp_neg(double t)
{
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    double kd = fma(-t, 1.4426950408889634, MAGIC);
    int k = lo_word(kd);
    double kf = kd - MAGIC;
    double r = fma(-kf, 6.93147180369123816490e-01, -t);     // ln2 split hi / lo (fdlibm)
    r = fma(-kf, 1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                       // 1/13!
    p = fma(p, r, 2.08767569878681e-09);                     // 1/12!
    p = fma(p, r, 2.505210838544172e-08);                    // 1/11!
    p = fma(p, r, 2.755731922398589e-07);                    // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);                   // 1/9!
    p = fma(p, r, 2.48015873015873e-05);                     // 1/8!
    p = fma(p, r, 1.984126984126984e-04);                    // 1/7!
    p = fma(p, r, 1.388888888888889e-03);                    // 1/6!
    p = fma(p, r, 8.333333333333333e-03);                    // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);                   // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);                   // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double y = from_words(hi_word(p) + (k << 20), lo_word(p));
    return (t > 707.0) ? 0.0 : y;
}

// cbrt(x) for positive normal x well inside the float range: seed r0 = x^(-1/3) from the single-precision
// MUFU pair lg2 / ex2 (relative error <~ 5e-6 over 1e-30 < x < 1e30), one quartically convergent FMA-only step
//   r1 = r0 (1 + e/3 + 2 e^2/9 + 14 e^3/81),  e = 1 - x r0^3      (truncation 35 e^4 / 243 < 1e-19)
// and cbrt(x) = x r1^2: ~15 instructions against ~40 for cbrt(); <= ~4 ulp.  Returns x^(1/3); rinv = x^(-1/3).
MK_HD double fast_cbrt_pos(double x, double& rinv)
{
    float xf = (float)x, lg, r0f;
#ifdef __CUDA_ARCH__
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(xf));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r0f) : "f"(lg * -0.33333334f));
#else
    lg = std::log2(xf);
    r0f = std::exp2(lg * -0.33333334f);
#endif
    double r0 = (double)r0f;
    double r2 = r0 * r0;
    double e = fma(-x * r0, r2, 1.0);
    double p = fma(fma(e, 14.0 / 81.0, 2.0 / 9.0), e, 1.0 / 3.0) * e;
    double r = fma(r0, p, r0);
    rinv = r;
    return (x * r) * r;
}

}  // namespace mk
