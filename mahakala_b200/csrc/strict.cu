// STRICT mode of the geodesic integrator: the reference's algorithm evaluated literally — forward-mode
// derivative of the metric (jets, the role of jax.jacfwd at geodesics.py:305), numerical 4x4 inverse
// (geodesics.py:347), the contraction of geodesics.py:307, RK4 as geodesics.py:317-336 and the step rule as
// geodesics.py:249-267 — in plain IEEE double arithmetic.  This file is compiled with --fmad=false, so no
// multiply-add is contracted and every +, -, *, /, sqrt rounds exactly as on the CPU; the results are
// BIT-IDENTICAL to the scalar C restatement that the tests use as their checker (they assert equality, not closeness).
// It exists to separate two questions: "is the GPU evaluating the same algorithm?" (strict mode, bit-exact)
// and "how far may the optimised closed-form kernel drift?" (fast mode, <= 1e-9).
// ~8x slower than the closed-form kernel; final-state outputs only.
#include "common.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {
namespace strict {

struct jet { double v, d[3]; };       // value + 3 spatial tangents (the t tangent is identically zero)

__device__ inline jet jconst(double c) { jet r; r.v = c; r.d[0] = r.d[1] = r.d[2] = 0.0; return r; }
__device__ inline jet jvar(double x, int k) { jet r = jconst(x); r.d[k] = 1.0; return r; }
__device__ inline jet jadd(jet a, jet b) { jet r; r.v = a.v + b.v; for (int k = 0; k < 3; k++) r.d[k] = a.d[k] + b.d[k]; return r; }
__device__ inline jet jsub(jet a, jet b) { jet r; r.v = a.v - b.v; for (int k = 0; k < 3; k++) r.d[k] = a.d[k] - b.d[k]; return r; }
__device__ inline jet jmul(jet a, jet b) { jet r; r.v = a.v * b.v; for (int k = 0; k < 3; k++) r.d[k] = a.d[k] * b.v + b.d[k] * a.v; return r; }
__device__ inline jet jdiv(jet a, jet b) { jet r; r.v = a.v / b.v; for (int k = 0; k < 3; k++) r.d[k] = (a.d[k] - b.d[k] * r.v) / b.v; return r; }
__device__ inline jet jscale(double c, jet a) { jet r; r.v = c * a.v; for (int k = 0; k < 3; k++) r.d[k] = c * a.d[k]; return r; }
__device__ inline jet jsqrt(jet a) { jet r; r.v = sqrt(a.v); for (int k = 0; k < 3; k++) r.d[k] = a.d[k] / (2.0 * r.v); return r; }

// geodesics.py:95-104 pushed through jets: g and dg/dx^(k+1)
__device__ void metric_and_jac(const double x[4], double a, double g[4][4], double jg[4][4][3])
{
    jet X = jvar(x[1], 0), Y = jvar(x[2], 1), Z = jvar(x[3], 2);
    double aa = a * a;
    jet zz = jmul(Z, Z);
    jet kk = jscale(0.5, jsub(jadd(jadd(jmul(X, X), jmul(Y, Y)), zz), jconst(aa)));
    jet rr = jadd(jsqrt(jadd(jmul(kk, kk), jscale(aa, zz))), kk);
    jet r = jsqrt(rr);
    jet f = jdiv(jmul(jscale(2.0, rr), r), jadd(jmul(rr, rr), jscale(aa, zz)));
    jet q = jadd(rr, jconst(aa));
    jet l[4];
    l[0] = jconst(1.0);
    l[1] = jdiv(jadd(jmul(r, X), jscale(a, Y)), q);
    l[2] = jdiv(jsub(jmul(r, Y), jscale(a, X)), q);
    l[3] = jdiv(Z, r);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            jet gij = jmul(f, jmul(l[i], l[j]));
            double eta = (i == j) ? ((i == 0) ? -1.0 : 1.0) : 0.0;
            g[i][j] = eta + gij.v;
            for (int k = 0; k < 3; k++) jg[i][j][k] = gij.d[k];
        }
}

// Gauss-Jordan with partial pivoting on [g | I]
__device__ void inv4(const double gin[4][4], double out[4][4])
{
    double m[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) { m[i][j] = gin[i][j]; m[i][j + 4] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int p = c;
        for (int r = c + 1; r < 4; r++) if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (p != c) for (int j = 0; j < 8; j++) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
        double piv = m[c][c];
        for (int j = 0; j < 8; j++) m[c][j] /= piv;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            double fct = m[r][c];
            for (int j = 0; j < 8; j++) m[r][j] -= fct * m[c][j];
        }
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[i][j] = m[i][j + 4];
}

// geodesics.py:294-309
__device__ void rhs(const double s[8], double a, double out[8])
{
    double g[4][4], jg[4][4][3], ig[4][4];
    const double* v = s + 4;
    metric_and_jac(s, a, g, jg);
    inv4(g, ig);
    double t1[4], t2[4] = {0, 0, 0, 0}, w[4];
    for (int i = 0; i < 4; i++) {
        double acc = 0;
        for (int j = 0; j < 4; j++) {
            double inner = 0;
            for (int k = 0; k < 3; k++) inner += jg[i][j][k] * v[k + 1];
            acc += inner * v[j];
        }
        t1[i] = acc;
    }
    for (int k = 0; k < 3; k++) {
        double acc = 0;
        for (int i = 0; i < 4; i++) {
            double inner = 0;
            for (int j = 0; j < 4; j++) inner += v[j] * jg[i][j][k];
            acc += v[i] * inner;
        }
        t2[k + 1] = acc;
    }
    for (int i = 0; i < 4; i++) w[i] = -t1[i] + 0.5 * t2[i];
    for (int i = 0; i < 4; i++) {
        double acc = 0;
        for (int j = 0; j < 4; j++) acc += ig[i][j] * w[j];
        out[4 + i] = acc;
        out[i] = v[i];
    }
}

// geodesics.py:317-336
__device__ void rk4(const double s[8], double dt, double a, double out[8])
{
    double k1[8], k2[8], k3[8], k4[8], tmp[8], r[8];
    rhs(s, a, r);   for (int i = 0; i < 8; i++) { k1[i] = dt * r[i]; tmp[i] = s[i] + 0.5 * k1[i]; }
    rhs(tmp, a, r); for (int i = 0; i < 8; i++) { k2[i] = dt * r[i]; tmp[i] = s[i] + 0.5 * k2[i]; }
    rhs(tmp, a, r); for (int i = 0; i < 8; i++) { k3[i] = dt * r[i]; tmp[i] = s[i] + k3[i]; }
    rhs(tmp, a, r); for (int i = 0; i < 8; i++) { k4[i] = dt * r[i]; }
    const double sixth = 1.0 / 6;
    for (int i = 0; i < 8; i++) out[i] = s[i] + sixth * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

// geodesics.py:284-291
__device__ double radius_cal(const double x[4], double a)
{
    double R = sqrt(x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
    double w = R * R - a * a;
    return sqrt((w + sqrt(w * w + 4 * (a * a) * (x[3] * x[3]))) / 2);
}

// geodesics.py:249-252
__device__ double step_rule(const double s[8], double div, double tol, double a, double rEH, double* r_out)
{
    double r = radius_cal(s, a);
    *r_out = r;
    double dt = -(r - rEH) / div;
    if (isnan(dt) || fabs(dt) * div < tol || fabs(dt) * div > 1500) return 0.0;
    return dt;
}

__global__ void __launch_bounds__(64) integrate_strict_kernel(const double* s0, long npx, int N, double div, double tol,
                                                              double a, double rEH, double* final_state, int* nsteps,
                                                              double* r_last, unsigned long long* total_steps)
{
    long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= npx) return;
    double s[8], cand[8];
    for (int i = 0; i < 8; i++) s[i] = s0[p * 8 + i];
    int n = 0;
    double r_cur, r_prev = 0.0, rl = 0.0;
    double best_dt = -1.0e300, r_before_best = 0.0;
    int best_idx = -1;
    double dt = step_rule(s, div, tol, a, rEH, &r_cur);
    bool terminated = false;
    for (int it = 0; it < N; it++) {
        double r_new = 0.0, dtn = 0.0;
        if (dt != 0.0) {
            rk4(s, dt, a, cand);
            dtn = step_rule(cand, div, tol, a, rEH, &r_new);
        }
        if (dt == 0.0 || dtn == 0.0) {
            if (best_dt > 0.0) rl = (best_idx >= 1) ? r_before_best : r_cur;
            else rl = (it >= 1) ? r_prev : r_cur;
            terminated = true;
            break;
        }
        if (dt > best_dt) { best_dt = dt; best_idx = it; r_before_best = r_prev; }
        r_prev = r_cur; r_cur = r_new;
        for (int i = 0; i < 8; i++) s[i] = cand[i];
        dt = dtn;
        n++;
    }
    if (!terminated) rl = (best_idx >= 1) ? r_before_best : r_prev;
    if (final_state) for (int i = 0; i < 8; i++) final_state[p * 8 + i] = s[i];
    if (nsteps) nsteps[p] = n;
    if (r_last) r_last[p] = rl;
    if (total_steps) atomicAdd(total_steps, (unsigned long long)n);
}

}  // namespace strict

int strict_integrate(double bhspin, long N, long npx, const double* s0, double div, double tol, double* final_state,
                     int* nsteps, double* r_last, unsigned long long* total_steps, cudaStream_t stream)
{
    double rEH = 1.0 + sqrt(1.0 - bhspin * bhspin);
    strict::integrate_strict_kernel<<<(unsigned)((npx + 63) / 64), 64, 0, stream>>>(s0, npx, (int)N, div, tol, bhspin, rEH,
                                                                                  final_state, nsteps, r_last, total_steps);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace mk
