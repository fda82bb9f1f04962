// Stand-alone camera kernels: initialize_geodesics_at_camera / get_camera_pixel / initial_condition
// (/root/reference/mahakala/geodesics.py:29-55, :107-134, :219-230).
#include "common.cuh"
#include "camera.cuh"
#include "integrate_kernel.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

__device__ __forceinline__ void store_state(double* s0, long idx, const double s[8])
{
    double4* p = reinterpret_cast<double4*>(s0 + idx * 8);
    p[0] = make_double4(s[0], s[1], s[2], s[3]);
    p[1] = make_double4(s[4], s[5], s[6], s[7]);
}

__global__ void camera_grid_kernel(KerrSchild g, CameraGeom c, double lo, double step, long n, int nullify, double* s0)
{
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    long ix = idx / n, iy = idx - ix * n;                   // meshgrid(indexing='ij').flatten()
    double x[4], v[4], s[8];
    camera_point(c, pixel_centre(lo, step, ix), pixel_centre(lo, step, iy), x, v);
    if (nullify) nullify_state(g, x, v, s);
    else { for (int m = 0; m < 4; m++) { s[m] = x[m]; s[4 + m] = v[m]; } }
    store_state(s0, idx, s);
}

__global__ void camera_points_kernel(KerrSchild g, CameraGeom c, const double* xi, const double* yi, long n, int nullify, double* s0)
{
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= n) return;
    double x[4], v[4], s[8];
    camera_point(c, xi[idx], yi[idx], x, v);
    if (nullify) nullify_state(g, x, v, s);
    else { for (int m = 0; m < 4; m++) { s[m] = x[m]; s[4 + m] = v[m]; } }
    store_state(s0, idx, s);
}

__global__ void initial_condition_kernel(KerrSchild g, const double* s0_x, const double* s0_v, long n, double* s0)
{
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= n) return;
    double x[4], v[4], s[8];
#pragma unroll
    for (int m = 0; m < 4; m++) { x[m] = s0_x[m * n + idx]; v[m] = s0_v[m * n + idx]; }
    nullify_state(g, x, v, s);
    store_state(s0, idx, s);
}

// find_shadow_bisection_angles (geodesics.py:405-435) in ONE launch: a lane owns an image-plane angle and repeats
// "ray through the mid radius -> integrate -> captured? -> halve the bracket" n_iter times.  The reference (and the
// host loop it replaces) launches one bundle per bisection iteration: 14 iterations x 4 kernels, each bounded by
// its longest ray.  Mid radius, image point and bracket update use non-contracted IEEE operations in the
// reference's order, so the radii are the ones the host loop returns.
__global__ void __launch_bounds__(32) shadow_bisection_kernel(KerrSchild g, CameraGeom cam, StepRule rule,
                                                              const double* __restrict__ cos_angle,
                                                              const double* __restrict__ sin_angle, long n, int N,
                                                              int n_iter, double inner0, double outer0, double limit,
                                                              double* __restrict__ inner_out, double* __restrict__ outer_out)
{
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= n) return;
    double inner = inner0, outer = outer0;
    const double ca = cos_angle[idx], sa = sin_angle[idx];
    for (int k = 0; k < n_iter; k++) {
        const double mid = __dadd_rn(__dmul_rn(__dsub_rn(outer, inner), 0.5), inner);     // (outer - inner) / 2 + inner
        double x[4], v[4], s[8];
        camera_point(cam, __dmul_rn(ca, mid), __dmul_rn(sa, mid), x, v);                   // geodesics.py:117-118
        nullify_state(g, x, v, s);
        int nsteps;
        const double r_last = integrate_one(g, rule, s, N, nsteps);
        if (r_last < limit) inner = mid;             // fell into the hole: the edge is further out
        else if (r_last >= limit) outer = mid;       // got away (a NaN radius moves neither end, as np.where)
    }
    inner_out[idx] = inner;
    outer_out[idx] = outer;
}

static KerrSchild make_ks(double a)
{
    KerrSchild g; g.set_spin(a);
    return g;
}

}  // namespace mk
using namespace mk;

extern "C" int mk_camera_grid(double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
                              double fov_upper, long n, int nullify, double* s0, void* stream)
{
    MK_REQUIRE(n >= 0, "pixels_per_side must be non-negative");
    if (n == 0) return 0;
    MK_REQUIRE(s0 != nullptr, "s0 is null");
    CameraGeom c = {cos_i, sin_i, distance};
    double step = (fov_upper - fov_lower) / (double)(2 * n);     // np.linspace: delta / div
    long total = n * n;
    camera_grid_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(make_ks(bhspin), c, fov_lower, step, n, nullify, s0);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_camera_points(double bhspin, double cos_i, double sin_i, double distance, const double* x_img,
                                const double* y_img, long n, int nullify, double* s0, void* stream)
{
    MK_REQUIRE(n >= 0, "n must be non-negative");
    if (n == 0) return 0;
    MK_REQUIRE(s0 && x_img && y_img, "null pointer");
    CameraGeom c = {cos_i, sin_i, distance};
    camera_points_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(make_ks(bhspin), c, x_img, y_img, n, nullify, s0);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_initial_condition(double bhspin, const double* s0_x, const double* s0_v, long n, double* s0, void* stream)
{
    MK_REQUIRE(n >= 0, "n must be non-negative");
    if (n == 0) return 0;
    MK_REQUIRE(s0 && s0_x && s0_v, "null pointer");
    initial_condition_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(make_ks(bhspin), s0_x, s0_v, n, s0);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_shadow_bisection(double bhspin, double cos_i, double sin_i, double distance, const double* cos_angle,
                                   const double* sin_angle, long n, long N, double div, double tol, int n_iter,
                                   double inner0, double outer0, double limit, double* inner_out, double* outer_out,
                                   void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(cos_angle && sin_angle && inner_out && outer_out, "null pointer");
    MK_REQUIRE(N >= 0 && N < (1L << 31) - 2 && n_iter >= 0, "N / n_iter out of range");
    MK_REQUIRE(div != 0.0, "div must be non-zero");
    KerrSchild g = make_ks(bhspin);
    CameraGeom c; c.ci = cos_i; c.si = sin_i; c.d = distance;
    StepRule rule;
    rule.div = div; rule.inv_div = 1.0 / div; rule.tol = tol; rule.rH = g.rH;
    // one warp per CTA: the CTAs spread over the SMs, so every warp runs at lone-warp latency (0.64 us per RK4 step)
    shadow_bisection_kernel<<<(unsigned)((n + 31) / 32), 32, 0, (cudaStream_t)stream>>>(g, c, rule, cos_angle, sin_angle, n, (int)N,
                                                                                      n_iter, inner0, outer0, limit,
                                                                                      inner_out, outer_out);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}
