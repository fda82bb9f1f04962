// Stand-alone camera kernels: initialize_geodesics_at_camera / get_camera_pixel / initial_condition
// (/root/reference/mahakala/geodesics.py:29-55, :107-134, :219-230).
#include "common.cuh"
#include "camera.cuh"
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
