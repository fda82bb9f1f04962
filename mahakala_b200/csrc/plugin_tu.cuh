// Translation-unit epilogue for RUN-TIME registered spacetimes (mk_register_metric): compiled by NVRTC after
// the user's source, which must define
//
//     struct UserMetric {
//         static constexpr bool stationary = true;    // OPTIONAL: metric independent of x[0] -> 3 tangents
//         double params[8];          // params[0] = the bhspin argument of the call; [1..7] = mk_metric_set_params
//         template <class T> __device__ void operator()(const T x[4], T g[4][4]) const;   // covariant metric
//         __device__ double radius(const double x[4]) const;     // radius used by the step rule
//         __device__ double horizon() const;                     // inner cut-off radius of the step rule
//     };
//
// (generic scalar T: double or mk::Dual<4>; use mk::mk_sqrt for square roots).  This mirrors how a user of the
// reference swaps spacetimes by replacing the module-level metric() (geodesics.py:88-104, :304-305); the
// derivative comes from forward-mode dual numbers exactly as jax.jacfwd provides it there.
#pragma once
#include "integrate_kernel.cuh"
#include "adaptive.cuh"
#include "metric_plugin.cuh"
#include "camera_nullify.cuh"
#include "render_kernel.cuh"

namespace mk {

struct PluginBlob {
    double params[8];
};

__device__ __forceinline__ DualMetric<UserMetric> plugin_metric(const PluginBlob& b)
{
    DualMetric<UserMetric> g;
#pragma unroll
    for (int i = 0; i < 8; i++) g.fn.params[i] = b.params[i];
    g.rH = g.fn.horizon();
    return g;
}

template <int MODE>
__device__ __forceinline__ void plugin_integrate(const PluginBlob& b, const IntegrateArgs& A0)
{
    DualMetric<UserMetric> g = plugin_metric(b);
    IntegrateArgs A = A0;
    A.rule.rH = g.rH;
    integrate_body<DualMetric<UserMetric>, MODE>(g, A);
}

}  // namespace mk

extern "C" __global__ void __launch_bounds__(128, 2) mk_plugin_integrate_final(mk::PluginBlob b, mk::IntegrateArgs A)
{
    mk::plugin_integrate<mk::MODE_FINAL>(b, A);
}
extern "C" __global__ void __launch_bounds__(128, 2) mk_plugin_integrate_padded(mk::PluginBlob b, mk::IntegrateArgs A)
{
    mk::plugin_integrate<mk::MODE_PADDED>(b, A);
}
extern "C" __global__ void __launch_bounds__(128, 2) mk_plugin_integrate_paged(mk::PluginBlob b, mk::IntegrateArgs A)
{
    mk::plugin_integrate<mk::MODE_PAGED>(b, A);
}

// optional adaptive integrator (adaptive.cuh) in the user's spacetime: the step follows the local error, not the
// Kerr-specific rule (r - r_H)/div; radius() and horizon() only decide when a ray has ended
extern "C" __global__ void __launch_bounds__(128, 2) mk_plugin_integrate_adaptive(mk::PluginBlob b, mk::AdaptiveArgs A)
{
    mk::DualMetric<UserMetric> g = mk::plugin_metric(b);
    A.rule.rH = g.rH;
    mk::integrate_adaptive_body(g, A);
}

// fused render (images.py:30-144) in the user's spacetime: geodesics from the dual-number plugin, fluid frame from
// its covariant / contravariant metric at every sample (athenak.py:760-786 with g, g^-1 of the plugin); one observing
// frequency per launch, any snapshot kind
extern "C" __global__ void __launch_bounds__(128, 2) mk_plugin_render(mk::PluginBlob b, mk::RenderArgs A)
{
    mk::DualMetric<UserMetric> g = mk::plugin_metric(b);
    A.rule.rH = g.rH;
    mk::render_body<mk::DualMetric<UserMetric>, 1, mk::SNAP_GENERIC>(g, A);
}

extern "C" __global__ void mk_plugin_rhs(mk::PluginBlob b, const double* state, long n, double* out)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    mk::DualMetric<UserMetric> g = mk::plugin_metric(b);
    double s[8], acc[4];
    for (int m = 0; m < 8; m++) s[m] = state[i * 8 + m];
    g.accel(s, s + 4, acc);
    for (int m = 0; m < 4; m++) { out[i * 8 + m] = s[4 + m]; out[i * 8 + 4 + m] = acc[m]; }
}

extern "C" __global__ void mk_plugin_rk4(mk::PluginBlob b, const double* state, const double* dt, long n, double* out)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    mk::DualMetric<UserMetric> g = mk::plugin_metric(b);
    double s[8], o[8];
    for (int m = 0; m < 8; m++) s[m] = state[i * 8 + m];
    mk::rk4_step(g, s, dt[i], o);
    for (int m = 0; m < 8; m++) out[i * 8 + m] = o[m];
}

extern "C" __global__ void mk_plugin_metric(mk::PluginBlob b, const double* x, long n, double* gout, double* giout)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    mk::DualMetric<UserMetric> g = mk::plugin_metric(b);
    double p[4], gc[4][4], gi[4][4];
    for (int m = 0; m < 4; m++) p[m] = x[i * 4 + m];
    g.metric_cov_con(p, gc, gi);
    for (int a = 0; a < 4; a++)
        for (int c = 0; c < 4; c++) {
            if (gout) gout[i * 16 + a * 4 + c] = gc[a][c];
            if (giout) giout[i * 16 + a * 4 + c] = gi[a][c];
        }
}

// initial_condition (geodesics.py:219-230) with the user's metric: s0_x, s0_v (4, n) -> s0 (n, 8)
extern "C" __global__ void mk_plugin_nullify(mk::PluginBlob b, const double* s0_x, const double* s0_v, long n, double* s0)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    mk::DualMetric<UserMetric> g = mk::plugin_metric(b);
    double x[4], v[4], gm[4][4], s[8];
    for (int m = 0; m < 4; m++) { x[m] = s0_x[m * n + i]; v[m] = s0_v[m * n + i]; }
    g.fn(x, gm);
    mk::nullify_with_metric(gm, x, v, s);
    for (int m = 0; m < 8; m++) s0[i * 8 + m] = s[m];
}
