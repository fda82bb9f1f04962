// Closed-form Cartesian Kerr-Schild metric plugin.
//
// Replaces, for the built-in spacetime, the reference's jacfwd(metric) + linalg.inv(metric) pair
// (/root/reference/mahakala/geodesics.py:88-104 metric, :294-309 rhs, :339-347 imetric):
//     g_mn = eta_mn + f l_m l_n ,  g^mn = eta^mn - f l^m l^n ,  l_m = (1, l1, l2, l3), l^m = (-1, l1, l2, l3)
// so the inverse is analytic and dg needs only grad f and grad l_i.  The geodesic acceleration
//     a^m = g^mn ( -d_k g_ns v^k v^s + 1/2 d_n g_ks v^k v^s )              (geodesics.py:307)
// is evaluated with shared sub-expressions; one reciprocal, one sqrt and one sqrt/rsqrt pair per call.
//
// Metric-plugin concept (see metric_plugin.cuh): a plugin provides
//     void   accel(const double x[4], const double v[4], double acc[4]) const;
//     double radius(const double x[4]) const;        // the step rule's radius (geodesics.py:284-291)
//     double horizon() const;                        // geodesics.py:350-351
//     void   metric_cov_con(const double x[4], double g[4][4], double gi[4][4]) const;   // sampling
#pragma once
#include "fp64_math.cuh"

namespace mk {

struct KerrSchild {
    static constexpr bool kHeavy = false;   // fits 4 CTAs of 128 threads per SM (126 registers)
    double a;    // spin
    double aa;   // a^2
    double rH;   // 1 + sqrt(1 - a^2), computed on the host exactly as geodesics.py:351
    double mhaa; // -a^2 / 2
    double a4;   // 4 a
    double maa2; // -2 a^2

    __host__ __device__ void set_spin(double spin)
    {
        a = spin; aa = spin * spin; rH = 1.0 + sqrt(1.0 - spin * spin);
        mhaa = -0.5 * aa; a4 = 4.0 * spin; maa2 = -2.0 * aa;
    }

    MK_HD double horizon() const { return rH; }

    // Point cache: the step rule's radius (geodesics.py:290-291, r = sqrt((w + sqrt(w^2 + 4 a^2 z^2))/2) with
    // w = R^2 - a^2) is the same quantity as the metric's r = sqrt(rr), rr = sqrt(kk^2 + a^2 z^2) + kk with
    // kk = w/2 (geodesics.py:99-101).  It is computed once per accepted state and reused by the first RK4
    // stage of the next step.
    struct Cache {
        double rr, r, ri;     // r^2, r, 1/r
    };

    MK_HD double radius(const double x[4], Cache& c) const
    {
        double zz = x[3] * x[3];
        double kk = fma(0.5, fma(x[1], x[1], fma(x[2], x[2], zz)), mhaa);
        c.rr = sqrt_only(fma(kk, kk, aa * zz)) + kk;
        pair_sqrt_rsqrt(c.rr, c.r, c.ri);
        return c.r;
    }

    MK_HD double radius(const double x[4]) const
    {
        Cache c;
        return radius(x, c);
    }

    // f and l_i of geodesics.py:97-103
    MK_HD void fl(const double x[4], double& f, double& l1, double& l2, double& l3) const
    {
        double zz = x[3] * x[3];
        double kk = 0.5 * (fma(x[1], x[1], fma(x[2], x[2], zz)) - aa);
        double az2 = aa * zz;
        double rr = fast_sqrt(fma(kk, kk, az2)) + kk;
        double r, ri;
        fast_sqrt_rsqrt(rr, r, ri);
        double den = fma(rr, rr, az2);
        double q = rr + aa;
        double inv = fast_rcp(den * q);
        double iden = inv * q, iq = inv * den;
        f = 2.0 * rr * r * iden;
        l1 = fma(r, x[1], a * x[2]) * iq;
        l2 = fma(r, x[2], -a * x[1]) * iq;
        l3 = x[3] * ri;
    }

    // same, reusing the point cache of x
    MK_HD void fl(const double x[4], const Cache& c, double& f, double& l1, double& l2, double& l3) const
    {
        double den = fma(c.rr, c.rr, aa * (x[3] * x[3]));
        double q = c.rr + aa;
        double inv = fast_rcp(den * q);
        double iden = inv * q, iq = inv * den;
        f = 2.0 * c.rr * c.r * iden;
        l1 = fma(c.r, x[1], a * x[2]) * iq;
        l2 = fma(c.r, x[2], -a * x[1]) * iq;
        l3 = x[3] * c.ri;
    }

    // Geodesic acceleration d v^m / d lambda (geodesics.py:301-309 in closed form).
    //
    // With g = eta + f l l (l_0 = 1, stationary) the lowered force is
    //     w_0 = -K ,   w_i = f L n_i + 1/2 L^2 d_i f - K l_i ,   K = (v.grad f) L + f M ,
    //     L = l_m v^m ,   n_i = (d_i l_j - d_j l_i) v^j ,   M = d_i l_j v^i v^j ,
    // and a^m = g^mn w_n with g^mn = eta^mn - f l^m l^n.  The spatial part of l is the principal null congruence of
    // Kerr seen in the flat background: geodesic and shear-free, so its gradient is pure expansion + twist,
    //     d_i l_j = t (delta_ij - l_i l_j) - omega eps_ijk l_k ,   t = r^3/den ,  omega = a z r/den ,  den = r^4 + a^2 z^2
    // (Re and Im of 1/(r + i a z/r)).  Hence
    //     n = 2 omega (l x v) ,   M = t (|v|^2 - (l.v)^2) ,   l.n = 0 ,   l.grad r = 1 ,
    // and the contraction P = l^n w_n = 1/2 L^2 (l.grad f) = 1/2 L^2 (alpha + beta l_3) needs no dot product with the
    // force either (grad f = alpha grad r + beta e_z).  87 FP64 operations per call (71 with the point cache)
    // against 103 for the component-wise form of round 1; identical to it up to rounding
    // (tests/test_host_harness_cpu.py holds both to the literal jets + 4x4 inverse of the CPU restatement).
    // The metric functions at the point of an acceleration call, for callers that need them as well (the fused
    // render kernel's fluid-frame algebra at the state it is about to step from).
    struct MetricFunctions {
        double f, l1, l2, l3;
    };

    MK_HD void accel(const double x[4], const double v[4], double acc[4], const Cache* cache = nullptr,
                     MetricFunctions* mf = nullptr) const
    {
        const double X = x[1], Y = x[2], Z = x[3];
        const double v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
        double zz = Z * Z;
        double az2 = aa * zz;
        double rr, r, ri;                                   // r^2, r, 1/r
        if (cache) {
            rr = cache->rr; r = cache->r; ri = cache->ri;
        } else {
            double kk = fma(0.5, fma(X, X, fma(Y, Y, zz)), mhaa);
            rr = sqrt_only(fma(kk, kk, az2)) + kk;
            pair_sqrt_rsqrt(rr, r, ri);
        }
        double den = fma(rr, rr, az2);                      // r^4 + a^2 z^2
        double q = rr + aa;
        double inv = fast_rcp(den * q);
        double iden = inv * q, iq = inv * den;              // 1/den, 1/q
        double u = rr * iden, rid = r * iden;
        double t = r * u;                                   // r^3 / den = f / 2 = expansion of l
        double l1 = fma(r, X, a * Y) * iq;
        double l2 = fma(r, Y, -(a * X)) * iq;
        double l3 = Z * ri;
        if (mf) { mf->f = t + t; mf->l1 = l1; mf->l2 = l2; mf->l3 = l3; }
        double zr = Z * rid;
        double raz = aa * zr;                               // grad r = (t x, t y, t z + raz)
        double om4 = a4 * zr;                               // 4 omega
        double Dr = fma(t, fma(X, v1, fma(Y, v2, Z * v3)), raz * v3);          // v . grad r
        double lv = fma(l1, v1, fma(l2, v2, l3 * v3));
        double L = lv + v0;                                 // l_m v^m
        double Mp = fma(-lv, lv, fma(v1, v1, fma(v2, v2, v3 * v3)));          // M / t
        // grad f = alpha grad r + beta e_z:  alpha / 2 = u (3 - 4 u r^2),  beta / 2 = -2 u raz
        double ha = u * fma(-4.0, u * rr, 3.0);
        double hb = u * (maa2 * zr);
        double Kh = fma(fma(hb, v3, ha * Dr), L, t * (t * Mp));               // K / 2
        double tL = t * L, L2 = L * L;
        double W = tL * om4;                                // f L 2 omega
        double cr1 = fma(l2, v3, -(l3 * v2)), cr2 = fma(l3, v1, -(l1 * v3)), cr3 = fma(l1, v2, -(l2 * v1));
        double haL2 = ha * L2, hbL2 = hb * L2;
        double hat = haL2 * t;
        // q_i = f L n_i + 1/2 L^2 d_i f
        double q1 = fma(W, cr1, hat * X);
        double q2 = fma(W, cr2, hat * Y);
        double q3 = fma(W, cr3, fma(hat, Z, fma(haL2, raz, hbL2)));
        double P = fma(hbL2, l3, haL2);                     // l^n w_n
        double a0h = fma(t, P, Kh);                         // (K + f P) / 2
        double a0 = a0h + a0h;
        acc[0] = a0;                                        // acc^0 = K + f P
        acc[1] = fma(-a0, l1, q1);                          // acc^i = w_i - f P l_i = q_i - (K + f P) l_i
        acc[2] = fma(-a0, l2, q2);
        acc[3] = fma(-a0, l3, q3);
    }

    // Round-1 form of the same acceleration (component-wise d_i l_j = g_i c_j + k_ij), kept as an independent
    // restatement for the host-side cross-check; no kernel calls it.
    MK_HD void accel_v1(const double x[4], const double v[4], double acc[4]) const
    {
        const double X = x[1], Y = x[2], Z = x[3];
        const double v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
        double zz = Z * Z;
        double az2 = aa * zz;
        double rr, r, ri;
        double kk = fma(0.5, fma(X, X, fma(Y, Y, zz)), -0.5 * aa);
        rr = quick_sqrt(fma(kk, kk, az2)) + kk;
        quick_sqrt_rsqrt(rr, r, ri);
        double den = fma(rr, rr, az2);
        double q = rr + aa;
        double inv = fast_rcp(den * q);
        double iden = inv * q, iq = inv * den;
        double rid = r * iden;
        double t = rid * rr;
        double f = t + t;
        double l1 = fma(r, X, a * Y) * iq;
        double l2 = fma(r, Y, -a * X) * iq;
        double l3 = Z * ri;
        double aaz = aa * Z;
        double g1 = t * X, g2 = t * Y, g3 = fma(t, Z, rid * aaz);
        double Dr = fma(v1, g1, fma(v2, g2, v3 * g3));
        double r2 = r + r;
        double c1 = fma(-r2, l1, X) * iq;
        double c2 = fma(-r2, l2, Y) * iq;
        double c3 = -(l3 * ri);
        double cv = fma(c1, v1, fma(c2, v2, c3 * v3));
        double p1 = v1 * iq, p2 = v2 * iq;
        double M = fma(Dr, cv, fma(r, fma(p1, v1, p2 * v2), (v3 * v3) * ri));
        double alpha = rr * iden * fma(-4.0 * f, r, 6.0);
        double beta = -2.0 * f * iden * aaz;
        double Df = fma(alpha, Dr, beta * v3);
        double L = fma(l1, v1, fma(l2, v2, fma(l3, v3, v0)));
        double K = fma(Df, L, f * M);
        double fL = f * L, hL2 = 0.5 * L * L;
        double ah = alpha * hL2;
        double A1 = fma(fL, cv, ah), B1 = fL * Dr, C1 = fL * (a + a);
        double q1 = fma(g1, A1, fma(-B1, c1, -C1 * p2));
        double q2 = fma(g2, A1, fma(-B1, c2, C1 * p1));
        double q3 = fma(g3, A1, fma(-B1, c3, hL2 * beta));
        double P = fma(l1, q1, fma(l2, q2, l3 * q3));
        double a0 = fma(f, P, K);
        acc[0] = a0;
        acc[1] = fma(-a0, l1, q1);
        acc[2] = fma(-a0, l2, q2);
        acc[3] = fma(-a0, l3, q3);
    }

    // Covariant and contravariant metric at x (for the fluid-frame algebra, athenak.py:760-762).
    MK_HD void metric_cov_con(const double x[4], double g[4][4], double gi[4][4]) const
    {
        double f, l[4];
        l[0] = 1.0;
        fl(x, f, l[1], l[2], l[3]);
        const double lu[4] = {-1.0, l[1], l[2], l[3]};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double eta = (i == j) ? (i == 0 ? -1.0 : 1.0) : 0.0;
                g[i][j] = fma(f * l[i], l[j], eta);
                gi[i][j] = fma(-f * lu[i], lu[j], eta);
            }
    }
};

}  // namespace mk
