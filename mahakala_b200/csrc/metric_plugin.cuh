// Metric-plugin interface of the ray kernels and the generic forward-mode (dual-number) adaptor.
//
// The reference has no explicit plugin API: rhs() calls the module-global metric()/imetric() and gets
// derivatives from jax.jacfwd (/root/reference/mahakala/geodesics.py:294-309, :339-347); a user swaps
// spacetimes by replacing those globals.  Here a spacetime is a C++ type with
//
//     struct Cache;                                                              // per-point scratch
//     void   accel(const double x[4], const double v[4], double acc[4], const Cache* = nullptr) const;
//     double radius(const double x[4]) const;  double radius(const double x[4], Cache&) const;   // step rule
//     double horizon() const;                                                    // inner cut-off radius
//     void   metric_cov_con(const double x[4], double g[4][4], double gi[4][4]) const;
//
// `KerrSchild` (ks_metric.cuh) implements it in closed form.  `DualMetric<Fn>` implements it for ANY
// user functor that can evaluate the covariant metric on a generic scalar type:
//
//     struct MyMetric {
//         template <class T> __device__ void operator()(const T x[4], T g[4][4]) const;  // fill all 16
//         __device__ double radius(const double x[4]) const;
//     };
//
// by pushing dual numbers (value + 4 tangents = jacfwd) through it and inverting the 4x4 metric with
// the adjugate (branch-free; no pivot problems at the ergosphere where g_tt changes sign).
#pragma once
#include "fp64_math.cuh"

namespace mk {

template <int NT>
struct Dual {
    double v;
    double d[NT];
    __device__ __forceinline__ Dual() {}
    __device__ __forceinline__ Dual(double c) : v(c)
    {
#pragma unroll
        for (int k = 0; k < NT; k++) d[k] = 0.0;
    }
};

#define MK_DUAL_FN template <int NT> __device__ __forceinline__ Dual<NT>

MK_DUAL_FN operator+(const Dual<NT>& a, const Dual<NT>& b)
{
    Dual<NT> r; r.v = a.v + b.v;
#pragma unroll
    for (int k = 0; k < NT; k++) r.d[k] = a.d[k] + b.d[k];
    return r;
}
MK_DUAL_FN operator-(const Dual<NT>& a, const Dual<NT>& b)
{
    Dual<NT> r; r.v = a.v - b.v;
#pragma unroll
    for (int k = 0; k < NT; k++) r.d[k] = a.d[k] - b.d[k];
    return r;
}
MK_DUAL_FN operator-(const Dual<NT>& a)
{
    Dual<NT> r; r.v = -a.v;
#pragma unroll
    for (int k = 0; k < NT; k++) r.d[k] = -a.d[k];
    return r;
}
MK_DUAL_FN operator*(const Dual<NT>& a, const Dual<NT>& b)
{
    Dual<NT> r; r.v = a.v * b.v;
#pragma unroll
    for (int k = 0; k < NT; k++) r.d[k] = fma(a.d[k], b.v, b.d[k] * a.v);
    return r;
}
MK_DUAL_FN operator/(const Dual<NT>& a, const Dual<NT>& b)
{
    Dual<NT> r;
    double ib = fast_rcp(b.v);
    r.v = a.v * ib;
#pragma unroll
    for (int k = 0; k < NT; k++) r.d[k] = fma(-b.d[k], r.v, a.d[k]) * ib;
    return r;
}
MK_DUAL_FN operator+(const Dual<NT>& a, double c) { Dual<NT> r = a; r.v += c; return r; }
MK_DUAL_FN operator+(double c, const Dual<NT>& a) { return a + c; }
MK_DUAL_FN operator-(const Dual<NT>& a, double c) { Dual<NT> r = a; r.v -= c; return r; }
MK_DUAL_FN operator-(double c, const Dual<NT>& a) { Dual<NT> r = -a; r.v += c; return r; }
MK_DUAL_FN operator*(const Dual<NT>& a, double c)
{
    Dual<NT> r; r.v = a.v * c;
#pragma unroll
    for (int k = 0; k < NT; k++) r.d[k] = a.d[k] * c;
    return r;
}
MK_DUAL_FN operator*(double c, const Dual<NT>& a) { return a * c; }
MK_DUAL_FN operator/(const Dual<NT>& a, double c) { return a * (1.0 / c); }
MK_DUAL_FN operator/(double c, const Dual<NT>& a) { return Dual<NT>(c) / a; }
MK_DUAL_FN dual_sqrt(const Dual<NT>& a)
{
    Dual<NT> r;
    double s, rs;
    fast_sqrt_rsqrt(a.v, s, rs);
    r.v = s;
    double h = 0.5 * rs;
#pragma unroll
    for (int k = 0; k < NT; k++) r.d[k] = a.d[k] * h;
    return r;
}
#undef MK_DUAL_FN

// plain-double overloads so that user functors can be instantiated with T = double as well
__device__ __forceinline__ double mk_sqrt(double x) { return fast_sqrt(x); }
template <int NT>
__device__ __forceinline__ Dual<NT> mk_sqrt(const Dual<NT>& x) { return dual_sqrt(x); }

// Inverse of a symmetric 4x4 matrix by the adjugate (2x2 sub-determinants), branch-free.
__device__ __forceinline__ void inverse4(const double m[4][4], double inv[4][4])
{
    double s0 = m[0][0] * m[1][1] - m[1][0] * m[0][1];
    double s1 = m[0][0] * m[1][2] - m[1][0] * m[0][2];
    double s2 = m[0][0] * m[1][3] - m[1][0] * m[0][3];
    double s3 = m[0][1] * m[1][2] - m[1][1] * m[0][2];
    double s4 = m[0][1] * m[1][3] - m[1][1] * m[0][3];
    double s5 = m[0][2] * m[1][3] - m[1][2] * m[0][3];
    double c5 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    double c4 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    double c3 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    double c2 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    double c1 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    double c0 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    double id = fast_rcp(det);
    inv[0][0] = ( m[1][1] * c5 - m[1][2] * c4 + m[1][3] * c3) * id;
    inv[0][1] = (-m[0][1] * c5 + m[0][2] * c4 - m[0][3] * c3) * id;
    inv[0][2] = ( m[3][1] * s5 - m[3][2] * s4 + m[3][3] * s3) * id;
    inv[0][3] = (-m[2][1] * s5 + m[2][2] * s4 - m[2][3] * s3) * id;
    inv[1][0] = (-m[1][0] * c5 + m[1][2] * c2 - m[1][3] * c1) * id;
    inv[1][1] = ( m[0][0] * c5 - m[0][2] * c2 + m[0][3] * c1) * id;
    inv[1][2] = (-m[3][0] * s5 + m[3][2] * s2 - m[3][3] * s1) * id;
    inv[1][3] = ( m[2][0] * s5 - m[2][2] * s2 + m[2][3] * s1) * id;
    inv[2][0] = ( m[1][0] * c4 - m[1][1] * c2 + m[1][3] * c0) * id;
    inv[2][1] = (-m[0][0] * c4 + m[0][1] * c2 - m[0][3] * c0) * id;
    inv[2][2] = ( m[3][0] * s4 - m[3][1] * s2 + m[3][3] * s0) * id;
    inv[2][3] = (-m[2][0] * s4 + m[2][1] * s2 - m[2][3] * s0) * id;
    inv[3][0] = (-m[1][0] * c3 + m[1][1] * c1 - m[1][2] * c0) * id;
    inv[3][1] = ( m[0][0] * c3 - m[0][1] * c1 + m[0][2] * c0) * id;
    inv[3][2] = (-m[3][0] * s3 + m[3][1] * s1 - m[3][2] * s0) * id;
    inv[3][3] = ( m[2][0] * s3 - m[2][1] * s1 + m[2][2] * s0) * id;
}

// A functor may declare `static constexpr bool stationary = true;` (no dependence on x[0]): the t tangent is
// then identically zero and only the 3 spatial tangents are propagated.
template <class T, class = void>
struct tangent_count {
    static constexpr int value = 4;
};
template <class T>
struct tangent_count<T, decltype((void)T::stationary)> {
    static constexpr int value = T::stationary ? 3 : 4;
};

// Generic plugin: derivatives by forward-mode duals through the user's metric functor.
template <class Fn>
struct DualMetric {
    static constexpr bool kHeavy = true;    // 16 metric jets live at once: give each thread the full register file
    Fn fn;
    double rH;

    struct Cache {};      // nothing worth carrying between the step rule and the next stage
    __device__ __forceinline__ double horizon() const { return rH; }
    __device__ __forceinline__ double radius(const double x[4]) const { return fn.radius(x); }
    __device__ __forceinline__ double radius(const double x[4], Cache&) const { return fn.radius(x); }

    __device__ __forceinline__ void metric_cov_con(const double x[4], double g[4][4], double gi[4][4]) const
    {
        fn(x, g);
        inverse4(g, gi);
    }

    // a^m = g^mn ( -d_k g_ns v^k v^s + 1/2 d_n g_ks v^k v^s )      (geodesics.py:307)
    __device__ __forceinline__ void accel(const double x[4], const double v[4], double acc[4],
                                          const Cache* = nullptr) const
    {
        constexpr int NT = tangent_count<Fn>::value;     // tangent k differentiates along x^(k + OFF)
        constexpr int OFF = 4 - NT;
        typedef Dual<NT> D;
        D xd[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            xd[m] = D(x[m]);
            if (m >= OFF) xd[m].d[m - OFF] = 1.0;
        }
        D gd[4][4];
        fn(xd, gd);
        double g[4][4], gi[4][4], w[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) g[i][j] = gd[i < j ? i : j][i < j ? j : i].v;      // symmetric: upper triangle
        inverse4(g, gi);
        // directional derivative along v of the (symmetric) metric, upper triangle only
        double Dg[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = i; j < 4; j++) {
                double dir = 0.0;
#pragma unroll
                for (int k = 0; k < NT; k++) dir = fma(gd[i][j].d[k], v[k + OFF], dir);
                Dg[i][j] = dir;
                Dg[j][i] = dir;
            }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double t1 = 0.0;        // sum_jk d_k g_ij v^k v^j
#pragma unroll
            for (int j = 0; j < 4; j++) t1 = fma(Dg[i][j], v[j], t1);
            double t2 = 0.0;        // sum_jk v^j v^k d_i g_jk  (zero for the t component of a stationary metric)
            if (i >= OFF) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double in = 0.5 * gd[j][j].d[i - OFF] * v[j];
#pragma unroll
                    for (int k = j + 1; k < 4; k++) in = fma(gd[j][k].d[i - OFF], v[k], in);
                    t2 = fma(in, v[j], t2);
                }
                t2 *= 2.0;
            }
            w[i] = fma(0.5, t2, -t1);
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            acc[i] = fma(gi[i][0], w[0], fma(gi[i][1], w[1], fma(gi[i][2], w[2], gi[i][3] * w[3])));
    }
};

// The Kerr-Schild metric typed generically (geodesics.py:95-104): exercises the dual-number path and
// serves as the template for user-registered spacetimes.
struct KerrSchildFn {
    static constexpr bool stationary = true;
    double a;

    template <class T>
    __device__ __forceinline__ void operator()(const T x[4], T g[4][4]) const
    {
        const double aa = a * a;
        T zz = x[3] * x[3];
        T kk = 0.5 * (x[1] * x[1] + x[2] * x[2] + zz - aa);
        T rr = mk_sqrt(kk * kk + aa * zz) + kk;
        T r = mk_sqrt(rr);
        T f = (2.0 * rr * r) / (rr * rr + aa * zz);
        T l[4];
        l[0] = T(1.0);
        l[1] = (r * x[1] + a * x[2]) / (rr + aa);
        l[2] = (r * x[2] - a * x[1]) / (rr + aa);
        l[3] = x[3] / r;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                T e = f * (l[i] * l[j]);
                g[i][j] = (i == j) ? e + (i == 0 ? -1.0 : 1.0) : e;
            }
    }

    __device__ __forceinline__ double radius(const double x[4]) const
    {
        const double aa = a * a;
        double R2 = fma(x[1], x[1], fma(x[2], x[2], x[3] * x[3]));
        double w = R2 - aa;
        double s = fast_sqrt(fma(w, w, 4.0 * aa * (x[3] * x[3])));
        return fast_sqrt(0.5 * (w + s));
    }
};

}  // namespace mk
