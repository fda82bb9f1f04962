// Geodesic integration kernels (integrate-only path): final state + classifier radius, and the
// trajectory-dump modes that reproduce the reference's stored (S, final_dt) outputs.  The kernel body
// lives in integrate_kernel.cuh; this file holds the __global__ instantiations for the built-in metric
// plugins and the C-ABI launchers.
#include "common.cuh"
#include "integrate_kernel.cuh"
#include "adaptive.cuh"
#include "ks_metric.cuh"
#include "metric_plugin.cuh"
#include "plugin.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

// Resident CTAs per SM, measured on B200 (scripts/variant_probe.py): the final/padded modes are fastest at
// 4 CTAs (124 registers), the paged dump at 3 (its page bookkeeping fits in registers and the scattered
// stores contend less): 25.9 -> 23.4 ms on cfg2.
#ifndef MK_PAGED_CTAS
#define MK_PAGED_CTAS 4
#endif
#ifndef MK_INT_CTAS           // experiment knob: resident CTAs of the final / padded modes (1 -> 255 registers)
#define MK_INT_CTAS 4
#endif
#ifndef MK_INT_THREADS        // experiment knob: threads per CTA (192 x 3 CTAs = 18 warps per SM at <= 112 registers)
#define MK_INT_THREADS 128
#endif
template <class Metric, int MODE, bool SHARED = false>
__global__ void __launch_bounds__(Metric::kHeavy ? 128 : MK_INT_THREADS, Metric::kHeavy ? 2 : ((MODE == MODE_PAGED) ? MK_PAGED_CTAS : MK_INT_CTAS)) integrate_kernel(const Metric g, const IntegrateArgs A)
{
    integrate_body<Metric, MODE, SHARED>(g, A);
}

// Rows after a ray's frozen row repeat the frozen state with dt = 0 (the reference's scan keeps
// emitting them, geodesics.py:264-269).  One thread per (row, ray).
__global__ void fill_frozen_rows_kernel(double* S, double* dt, const double* final_state,
                                        const int32_t* nsteps, long npx, long nrows)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long total = npx * nrows;
    for (; i < total; i += (long)gridDim.x * blockDim.x) {
        long row = i / npx, ray = i - row * npx;
        if (row > nsteps[ray]) {
            const double4* f = reinterpret_cast<const double4*>(final_state + ray * 8);
            double4* p = reinterpret_cast<double4*>(S + i * 8);
            p[0] = f[0];
            p[1] = f[1];
            dt[i] = 0.0;
        }
    }
}

// N == 0: the scan has no iterations (geodesics.py:272).  Every ray "ends" where it started: final = s0, no steps,
// classifier radius = the Kerr-Schild radius of s0.
__global__ void zero_steps_kernel(KerrSchild g, const double* __restrict__ s0, long npx, double* final_state,
                                  int32_t* nsteps, double* r_last)
{
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < npx; p += (long)gridDim.x * blockDim.x) {
        double s[8];
#pragma unroll
        for (int m = 0; m < 8; m++) s[m] = s0[p * 8 + m];
        if (final_state)
#pragma unroll
            for (int m = 0; m < 8; m++) final_state[p * 8 + m] = s[m];
        if (nsteps) nsteps[p] = 0;
        if (r_last) r_last[p] = g.radius(s);
    }
}

static int launch_zero_steps(double bhspin, const double* s0, long npx, double* final_state, int32_t* nsteps,
                             double* r_last, cudaStream_t stream)
{
    if (npx == 0 || (!final_state && !nsteps && !r_last)) return 0;
    KerrSchild g; g.set_spin(bhspin);
    long blocks = (npx + 127) / 128, cap = (long)sm_count() * 16;
    zero_steps_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, stream>>>(g, s0, npx, final_state, nsteps, r_last);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int strict_integrate(double bhspin, long N, long npx, const double* s0, double div, double tol, double* final_state,
                     int* nsteps, double* r_last, unsigned long long* total_steps, cudaStream_t stream);

template <class Metric>
static int launch_integrate(const Metric& g, const IntegrateArgs& A, cudaStream_t stream)
{
    int per_sm = 0;
    const bool shared = A.chunk_div > 0;       // set by mk_integrate_shared (number of GPUs on the queue)
    auto kern = A.pages ? (shared ? integrate_kernel<Metric, MODE_PAGED, true> : integrate_kernel<Metric, MODE_PAGED>)
                        : (A.S ? integrate_kernel<Metric, MODE_PADDED> : integrate_kernel<Metric, MODE_FINAL>);
    const size_t dyn_smem = (A.pages && MK_DUMP_TMA) ? 4 * DUMP_SMEM_PER_WARP : 0;      // TMA dump staging, 4 warps
    const int threads = Metric::kHeavy ? 128 : MK_INT_THREADS;
    const int wpc = threads / 32;                                            // warps per CTA
    MK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, dyn_smem));
    if (per_sm < 1) per_sm = 1;
    long warps_needed = (A.npx + 31) / 32;
    long blocks = (long)sm_count() * per_sm;
    long need = (warps_needed + wpc - 1) / wpc;
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    IntegrateArgs B = A;
    B.chunk_div = (int)(4 * wpc * blocks) * (shared ? A.chunk_div : 1);     // 4 x warps x participating GPUs
    B.chunk_mul = (unsigned)(0x100000000ULL / (unsigned long long)B.chunk_div);
    kern<<<(unsigned)blocks, threads, dyn_smem, stream>>>(g, B);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace mk

using namespace mk;

static int dispatch_integrate(int metric_id, double bhspin, IntegrateArgs& A, cudaStream_t stream);

extern "C" int mk_integrate(int metric_id, double bhspin, long N, long npx, const double* s0, double div,
                            double tol, double* final_state, int32_t* nsteps, double* r_last, double* S,
                            double* dt, long nrows, unsigned long long* total_steps, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MK_REQUIRE(npx >= 0 && N >= 0, "npx and N must be non-negative");
    MK_REQUIRE(N < (1L << 31) - 2, "N too large");
    MK_REQUIRE((S == nullptr) == (dt == nullptr), "S and dt dump buffers must be given together");
    MK_REQUIRE(npx == 0 || s0 != nullptr, "s0 is null");
    MK_REQUIRE(div != 0.0, "div must be non-zero");
    if (npx == 0) return 0;
    if (N == 0) return launch_zero_steps(bhspin, s0, npx, final_state, nsteps, r_last, stream);
    IntegrateArgs A;
    A.s0 = s0; A.npx = npx; A.N = (int)N;
    A.rule.div = div; A.rule.inv_div = 1.0 / div; A.rule.tol = tol;
    A.final_state = final_state; A.nsteps = nsteps; A.r_last = r_last;
    A.S = S; A.dt = dt; A.nrows = S ? nrows : 0;
    A.queue = queue_counter(stream, 0);
    if (!A.queue) return 1;
    A.ray_order = nullptr; A.page_id_offset = 0; A.chunk_div = 0; A.chunk_mul = 0;
    A.total_steps = total_steps;
    A.pages = nullptr; A.page_next = nullptr; A.page_first = nullptr; A.page_counter = nullptr;
    A.max_pages = 0; A.overflow = nullptr;
    return dispatch_integrate(metric_id, bhspin, A, stream);
}

static int dispatch_integrate(int metric_id, double bhspin, IntegrateArgs& A, cudaStream_t stream)
{
    int rc;
    if (metric_id == MK_METRIC_KERR_SCHILD) {
        KerrSchild g; g.set_spin(bhspin);
        A.rule.rH = g.rH;
        rc = launch_integrate(g, A, stream);
    } else if (metric_id == MK_METRIC_KERR_SCHILD_DUAL) {
        DualMetric<KerrSchildFn> g; g.fn.a = bhspin; g.rH = 1.0 + sqrt(1.0 - bhspin * bhspin);
        A.rule.rH = g.rH;
        rc = launch_integrate(g, A, stream);
    } else if (metric_id == MK_METRIC_KERR_SCHILD_STRICT) {
        if (A.S || A.pages) {
            set_error("the strict (literal IEEE) integrator provides final states only; use a dump mode of the default metric");
            return 2;
        }
        rc = strict_integrate(bhspin, A.N, A.npx, A.s0, A.rule.div, A.rule.tol, A.final_state, A.nsteps, A.r_last,
                              A.total_steps, stream);
    } else if (metric_id >= MK_METRIC_PLUGIN_BASE) {
        rc = plugin_integrate(metric_id, bhspin, A, stream);
    } else {
        set_error("unknown metric id %d", metric_id);
        return 2;
    }
    return rc;
}

// ---- optional adaptive integrator (embedded Dormand-Prince 5(4), adaptive.cuh) -------------------------------------
template <class Metric>
__global__ void __launch_bounds__(128) integrate_adaptive_kernel(const Metric g, const AdaptiveArgs A)
{
    integrate_adaptive_body(g, A);
}

template <class Metric>
static int launch_adaptive(const Metric& g, const AdaptiveArgs& A, cudaStream_t stream)
{
    long blocks = (A.npx + 127) / 128;
    const long cap = (long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    integrate_adaptive_kernel<Metric><<<(unsigned)blocks, 128, 0, stream>>>(g, A);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_integrate_adaptive(int metric_id, double bhspin, long N, long npx, const double* s0, double rtol,
                                     double atol, double tol, double cap, double* final_state, int32_t* nsteps,
                                     int32_t* nrejected, double* r_last, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MK_REQUIRE(npx >= 0 && N >= 0 && N < (1L << 31) - 2, "npx / N out of range");
    MK_REQUIRE(npx == 0 || s0 != nullptr, "s0 is null");
    MK_REQUIRE(rtol > 0.0 && atol >= 0.0, "rtol must be positive, atol non-negative");
    MK_REQUIRE(cap > 0.0 && cap < 1.0, "cap must be in (0, 1): a step never reaches the horizon");
    if (npx == 0) return 0;
    AdaptiveArgs A;
    A.s0 = s0; A.npx = npx; A.N = (int)N;
    A.rule.rtol = rtol; A.rule.atol = atol; A.rule.tol = tol; A.rule.far = 1500.0; A.rule.cap = cap; A.rule.div0 = 40.0;
    A.final_state = final_state; A.nsteps = nsteps; A.nrejected = nrejected; A.r_last = r_last;
    if (metric_id == MK_METRIC_KERR_SCHILD) {
        KerrSchild g; g.set_spin(bhspin);
        A.rule.rH = g.rH;
        return launch_adaptive(g, A, stream);
    }
    if (metric_id == MK_METRIC_KERR_SCHILD_DUAL) {
        DualMetric<KerrSchildFn> g; g.fn.a = bhspin; g.rH = 1.0 + sqrt(1.0 - bhspin * bhspin);
        A.rule.rH = g.rH;
        return launch_adaptive(g, A, stream);
    }
    if (metric_id >= MK_METRIC_PLUGIN_BASE) {
        A.rule.rH = 0.0;        // set from the plugin's horizon() in the kernel
        void* extra[] = {&A};
        return plugin_elementwise(metric_id, bhspin, "mk_plugin_integrate_adaptive", extra, 1, npx, stream);
    }
    set_error("the adaptive integrator runs with the Kerr-Schild metrics or a registered one (metric id %d)", metric_id);
    return 2;
}

extern "C" int mk_fill_frozen_rows(double* S, double* dt, const double* final_state, const int32_t* nsteps,
                                   long npx, long nrows, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (npx * nrows == 0) return 0;
    long total = npx * nrows;
    long blocks = (total + 255) / 256;
    long cap = (long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    fill_frozen_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(S, dt, final_state, nsteps, npx, nrows);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- paged (single-pass, ragged) trajectory dump ------------------------------------------------------
extern "C" int mk_page_rows(void) { return PAGE_SLOTS; }

static int integrate_paged_impl(int metric_id, double bhspin, long N, long npx, const double* s0, double div,
                                double tol, double* final_state, int32_t* nsteps, double* r_last, double* pages,
                                int32_t* page_next, int32_t* page_first, unsigned int* page_counter, long max_pages,
                                int32_t* overflow, unsigned long long* total_steps, unsigned int* queue,
                                const int32_t* ray_order, long page_id_offset, int participants, cudaStream_t stream)
{
    MK_REQUIRE(npx >= 0 && N >= 0 && N < (1L << 31) - 2, "npx / N out of range");
    MK_REQUIRE(npx < (1L << 31) - 64, "more than 2^31 rays per launch");
    MK_REQUIRE(pages && page_next && page_first && page_counter && overflow, "null page-store pointer");
    MK_REQUIRE(max_pages > 0 && max_pages < (1L << 31), "max_pages out of range");
    MK_REQUIRE(page_id_offset >= 0 && page_id_offset + max_pages < (1L << 31), "page_id_offset out of range");
    MK_REQUIRE(npx == 0 || s0 != nullptr, "s0 is null");
    MK_REQUIRE(div != 0.0, "div must be non-zero");
    if (npx == 0) return 0;
    if (N == 0) {                                          // no rows, no pages
        MK_REQUIRE(!queue && !ray_order, "N == 0 is not supported with a shared queue");
        return launch_zero_steps(bhspin, s0, npx, final_state, nsteps, r_last, stream);
    }
    IntegrateArgs A;
    A.s0 = s0; A.npx = npx; A.N = (int)N;
    A.rule.div = div; A.rule.inv_div = 1.0 / div; A.rule.tol = tol;
    A.final_state = final_state; A.nsteps = nsteps; A.r_last = r_last;
    A.S = nullptr; A.dt = nullptr; A.nrows = 0;
    A.queue = queue ? queue : queue_counter(stream, 0);
    if (!A.queue) return 1;
    A.ray_order = ray_order; A.page_id_offset = (int)page_id_offset;
    // > 0 selects the shared-queue kernel variant; launch_integrate multiplies by 4 x its own warps
    A.chunk_div = queue ? (participants > 0 ? participants : 1) : 0;
    A.chunk_mul = 0;
    A.total_steps = total_steps;
    A.pages = pages; A.page_next = page_next; A.page_first = page_first; A.page_counter = page_counter;
    A.max_pages = (unsigned)max_pages; A.overflow = overflow;
    return dispatch_integrate(metric_id, bhspin, A, stream);
}

extern "C" int mk_integrate_paged(int metric_id, double bhspin, long N, long npx, const double* s0, double div,
                                  double tol, double* final_state, int32_t* nsteps, double* r_last,
                                  double* pages, int32_t* page_next, int32_t* page_first,
                                  unsigned int* page_counter, long max_pages, int32_t* overflow,
                                  unsigned long long* total_steps, void* stream_)
{
    return integrate_paged_impl(metric_id, bhspin, N, npx, s0, div, tol, final_state, nsteps, r_last, pages, page_next,
                                page_first, page_counter, max_pages, overflow, total_steps, nullptr, nullptr, 0, 1,
                                (cudaStream_t)stream_);
}

extern "C" int mk_integrate_shared(int metric_id, double bhspin, long N, long npx, const double* s0, double div,
                                   double tol, double* final_state, int32_t* nsteps, double* r_last,
                                   double* pages, int32_t* page_next, int32_t* page_first,
                                   unsigned int* page_counter, long max_pages, int32_t* overflow,
                                   unsigned long long* total_steps, unsigned int* queue, const int32_t* ray_order,
                                   long page_id_offset, int participants, void* stream_)
{
    MK_REQUIRE(queue != nullptr, "queue is null (use mk_integrate_paged for a private queue)");
    return integrate_paged_impl(metric_id, bhspin, N, npx, s0, div, tol, final_state, nsteps, r_last, pages, page_next,
                                page_first, page_counter, max_pages, overflow, total_steps, queue, ray_order,
                                page_id_offset, participants, (cudaStream_t)stream_);
}

namespace mk {
// paged store -> the reference's padded layout for a selection of rays: S (nrows, nsel, 8), dt (nrows, nsel).
// One thread per selected ray walks the column of its warp's log; rows past the ray's last stored row repeat
// it with dt = 0.
__global__ void paged_gather_kernel(const double* __restrict__ pages, const int* __restrict__ page_next,
                                    const int* __restrict__ page_first, const int32_t* __restrict__ nsteps,
                                    const long* __restrict__ ray_idx, long nsel, long nrows, long N,
                                    double* __restrict__ S, double* __restrict__ dt)
{
    long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j >= nsel) return;
    long ray = ray_idx ? ray_idx[j] : j;
    long have = (long)nsteps[ray] + 1;          // rows stored: 0..n (frozen row) or N rows when n == N
    if (have > N) have = N;
    int page = page_first[2 * ray];
    int sl = page_first[2 * ray + 1];
    int slot = sl >> 5;
    const int ln = sl & 31;
    double4 lo = make_double4(0, 0, 0, 0), hi = lo;
    for (long row = 0; row < nrows; row++) {
        double d = 0.0;
        if (row < have && page >= 0) {
            const double* pg = pages + (long)page * PAGE_DOUBLES;
            const int rr = slot * 32 + ln;
            const double4* p = reinterpret_cast<const double4*>(pg + rr * 8);
            lo = p[0]; hi = p[1];
            d = pg[PAGE_SLOTS * 32 * 8 + rr];
            if (++slot == PAGE_SLOTS) { slot = 0; page = page_next[page]; }
        }
        double4* o = reinterpret_cast<double4*>(S + (row * nsel + j) * 8);
        o[0] = lo; o[1] = hi;
        dt[row * nsel + j] = d;
    }
}
}  // namespace mk

extern "C" int mk_paged_gather(const double* pages, const int32_t* page_next, const int32_t* page_first,
                               const int32_t* nsteps, const long* ray_idx, long nsel, long nrows, long N,
                               double* S, double* dt, void* stream)
{
    if (nsel <= 0 || nrows <= 0) return 0;
    MK_REQUIRE(pages && page_next && page_first && nsteps && S && dt, "null pointer");
    paged_gather_kernel<<<(unsigned)((nsel + 63) / 64), 64, 0, (cudaStream_t)stream>>>(pages, page_next, page_first, nsteps,
                                                                                    ray_idx, nsel, nrows, N, S, dt);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}
