// Fused render kernel: camera ray -> RK4 geodesic -> snapshot sample -> j_nu, alpha_nu -> intensity.
// The kernel body (generic in the spacetime) lives in render_kernel.cuh; this file holds the __global__
// instantiations for the built-in Kerr-Schild metric and the C-ABI launchers.
#include <cstdlib>
#include "common.cuh"
#include "render_kernel.cuh"
#include "snapshot.cuh"
#include "plugin.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

template <int NF, int KIND>
__global__ void MK_RENDER_BOUNDS render_kernel(const KerrSchild g, const RenderArgs A)
{
    render_body<KerrSchild, NF, KIND>(g, A);
}

// long-patch variant (render_long.cu): producer warp (geodesic) + consumer warps (sample, emission)
int launch_render_pipeline(const KerrSchild& g, const RenderArgs& A, int nfreq, long npatches, int exclusive, int max_ctas,
                           cudaStream_t stream);

// (Round 1 kept an experimental lane-refill variant of this kernel here: incoherent warps lose the L1 locality of
// the gathers, 1.25-1.95x slower than whole patches on cfg4; numbers in DESIGN.md.)

template <int NF, int KIND>
static int launch_render_kind(const KerrSchild& g, const RenderArgs& A, long npatches, cudaStream_t stream);

template <int NF>
static int launch_render(const KerrSchild& g, const RenderArgs& A, long npatches, cudaStream_t stream)
{
    switch (snapshot_kind(A.sn)) {
        case SNAP_F64_GRID_POW2: return launch_render_kind<NF, SNAP_F64_GRID_POW2>(g, A, npatches, stream);
        case SNAP_F32_GRID_POW2: return launch_render_kind<NF, SNAP_F32_GRID_POW2>(g, A, npatches, stream);
        default: return launch_render_kind<NF, SNAP_GENERIC>(g, A, npatches, stream);
    }
}

template <int NF, int KIND>
static int launch_render_kind(const KerrSchild& g, const RenderArgs& A, long npatches, cudaStream_t stream)
{
#ifdef MK_RENDER_SMEM_STAGE
    const size_t dyn_smem = (MK_RENDER_THREADS / 32) * (1728 + 8);
#else
    const size_t dyn_smem = 0;
#endif
    int per_sm = 0;
    MK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel<NF, KIND>, MK_RENDER_THREADS, dyn_smem));
    if (per_sm < 1) per_sm = 1;
    long blocks = (long)sm_count() * per_sm;
    const long warps = MK_RENDER_THREADS / 32;
    long need = (npatches + warps - 1) / warps;
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    render_kernel<NF, KIND><<<(unsigned)blocks, MK_RENDER_THREADS, dyn_smem, stream>>>(g, A);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace mk
using namespace mk;

extern "C" long mk_render_patch_count(long res, const double* s0, long npx)
{
    if (s0) return (npx + 31) / 32;
    return ((res + PATCH_X - 1) / PATCH_X) * ((res + PATCH_Y - 1) / PATCH_Y);
}

static int render_impl(int metric_id, double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
                       double fov_upper, long res, const double* s0, long npx, long N, double div,
                       double tol, const mk_snapshot* snap, const mk_emission_params* params, int nfreq,
                       const double* nu_obs, double* image, int32_t* nsteps,
                       unsigned long long* total_steps, unsigned long long* total_samples,
                       unsigned int* queue, long patch_begin, long patch_end, long patch_stride,
                       const int* patch_order, cudaStream_t stream, int pipeline = 0, int exclusive = 0, int max_ctas = 0)
{
    MK_REQUIRE(snap && params && nu_obs && image, "null pointer");
    MK_REQUIRE(nfreq >= 1 && nfreq <= 8, "nfreq must be in 1..8");
    MK_REQUIRE(N >= 0 && N < (1L << 31) - 2, "N out of range");
    MK_REQUIRE(div != 0.0, "div must be non-zero");
    if (!s0) { MK_REQUIRE(res >= 0 && npx == res * res, "npx must equal res*res for the grid camera"); }
    if (npx == 0) return 0;
    KerrSchild g;
    g.set_spin(bhspin);
    RenderArgs A;
    A.cam.ci = cos_i; A.cam.si = sin_i; A.cam.d = distance;
    A.fov_lo = fov_lower;
    A.step = res > 0 ? (fov_upper - fov_lower) / (double)(2 * res) : 0.0;
    A.res = res;
    A.patches_y = (res + PATCH_Y - 1) / PATCH_Y;
    A.s0 = s0; A.npx = npx; A.N = (int)N;
    A.rule.div = div; A.rule.inv_div = 1.0 / div; A.rule.tol = tol; A.rule.rH = g.rH;
    A.sn = snap->view;
    memcpy(&A.P, params, sizeof A.P);
    A.C = make_emission_consts(A.P, nu_obs, nfreq);
    for (int f = 0; f < 8; f++) {
        A.nu_obs[f] = nu_obs[f < nfreq ? f : nfreq - 1];
        A.inv_nu_obs[f] = 1.0 / A.nu_obs[f];
    }
    A.image = image; A.nsteps = nsteps; A.total_steps = total_steps; A.total_samples = total_samples;
    long npatches = mk_render_patch_count(res, s0, npx);
    A.patch_begin = patch_begin < 0 ? 0 : patch_begin;
    A.patch_end = (patch_end < 0 || patch_end > npatches) ? npatches : patch_end;
    A.patch_stride = patch_stride < 1 ? 1 : patch_stride;
    A.patch_order = patch_order;
    A.pipe_groups = 1;
    if (A.patch_begin >= A.patch_end) return 0;
    long span = (A.patch_end - A.patch_begin + A.patch_stride - 1) / A.patch_stride;
    if (metric_id >= MK_METRIC_PLUGIN_BASE) {
        // run-time registered spacetime: the NVRTC-built kernel carries one frequency; more are separate launches,
        // each with its own queue counter (a caller-supplied shared queue serves exactly one launch)
        MK_REQUIRE(queue == nullptr || nfreq == 1, "a shared queue serves one frequency per launch with a registered spacetime");
        for (int f = 0; f < nfreq; f++) {
            RenderArgs B = A;
            B.C = make_emission_consts(A.P, nu_obs + f, 1);
            B.nu_obs[0] = nu_obs[f]; B.inv_nu_obs[0] = 1.0 / nu_obs[f];
            B.image = image + (long)f * npx;
            if (f > 0) { B.nsteps = nullptr; B.total_steps = nullptr; B.total_samples = nullptr; }
            B.queue = queue ? queue : queue_counter(stream, 1);
            if (!B.queue) return 1;
            if (int rc = plugin_render(metric_id, bhspin, &B, sizeof B, span, stream)) return rc;
        }
        return 0;
    }
    MK_REQUIRE(metric_id == MK_METRIC_KERR_SCHILD || metric_id == MK_METRIC_KERR_SCHILD_DUAL,
               "the fused render runs with the built-in Kerr-Schild spacetime or a registered one");
    A.queue = queue ? queue : queue_counter(stream, 1);
    if (!A.queue) return 1;
    if (pipeline) {
        MK_REQUIRE(metric_id == MK_METRIC_KERR_SCHILD, "the long-patch pipeline runs in the built-in Kerr-Schild spacetime");
        return launch_render_pipeline(g, A, nfreq, span, exclusive, max_ctas, stream);
    }
    if (nfreq == 1) return launch_render<1>(g, A, span, stream);
    if (nfreq == 2) return launch_render<2>(g, A, span, stream);
    if (nfreq == 3) return launch_render<3>(g, A, span, stream);
    if (nfreq == 4) return launch_render<4>(g, A, span, stream);
    if (nfreq == 5) return launch_render<5>(g, A, span, stream);
    if (nfreq == 6) return launch_render<6>(g, A, span, stream);
    if (nfreq == 7) return launch_render<7>(g, A, span, stream);
    return launch_render<8>(g, A, span, stream);
}

extern "C" int mk_render(double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
                         double fov_upper, long res, const double* s0, long npx, long N, double div,
                         double tol, const mk_snapshot* snap, const mk_emission_params* params, int nfreq,
                         const double* nu_obs, double* image, int32_t* nsteps,
                         unsigned long long* total_steps, unsigned long long* total_samples,
                         unsigned int* queue, long patch_begin, long patch_end, long patch_stride,
                         const int* patch_order, void* stream_)
{
    return render_impl(MK_METRIC_KERR_SCHILD, bhspin, cos_i, sin_i, distance, fov_lower, fov_upper, res, s0, npx, N, div,
                       tol, snap, params, nfreq, nu_obs, image, nsteps, total_steps, total_samples, queue, patch_begin,
                       patch_end, patch_stride, patch_order, (cudaStream_t)stream_);
}

extern "C" int mk_render_long(double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
                              double fov_upper, long res, const double* s0, long npx, long N, double div,
                              double tol, const mk_snapshot* snap, const mk_emission_params* params, int nfreq,
                              const double* nu_obs, double* image, int32_t* nsteps,
                              unsigned long long* total_steps, unsigned long long* total_samples,
                              unsigned int* queue, long patch_begin, long patch_end, long patch_stride,
                              const int* patch_order, int exclusive, int max_ctas, void* stream_)
{
    return render_impl(MK_METRIC_KERR_SCHILD, bhspin, cos_i, sin_i, distance, fov_lower, fov_upper, res, s0, npx, N, div,
                       tol, snap, params, nfreq, nu_obs, image, nsteps, total_steps, total_samples, queue, patch_begin,
                       patch_end, patch_stride, patch_order, (cudaStream_t)stream_, 1, exclusive, max_ctas);
}

extern "C" int mk_render_metric(int metric_id, double bhspin, double cos_i, double sin_i, double distance,
                                double fov_lower, double fov_upper, long res, const double* s0, long npx, long N,
                                double div, double tol, const mk_snapshot* snap, const mk_emission_params* params,
                                int nfreq, const double* nu_obs, double* image, int32_t* nsteps,
                                unsigned long long* total_steps, unsigned long long* total_samples,
                                unsigned int* queue, long patch_begin, long patch_end, long patch_stride,
                                const int* patch_order, void* stream_)
{
    return render_impl(metric_id, bhspin, cos_i, sin_i, distance, fov_lower, fov_upper, res, s0, npx, N, div, tol, snap,
                       params, nfreq, nu_obs, image, nsteps, total_steps, total_samples, queue, patch_begin, patch_end,
                       patch_stride, patch_order, (cudaStream_t)stream_);
}
