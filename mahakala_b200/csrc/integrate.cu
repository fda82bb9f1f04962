// Geodesic integration kernels (integrate-only path): final state + classifier radius, and the
// trajectory-dump mode that reproduces the reference's stored (S, final_dt) outputs.
//
// Replaces /root/reference/mahakala/geodesics.py:233-281 (geodesic_integrator) and the last-point
// rule of :370-378.  One ray per lane, state in registers.  The kernel is persistent: each warp pulls
// rays from a global queue and REFILLS lanes whose ray has frozen (ballot + one atomic per refill), so
// warps stay full although step counts vary ~5x across the image (photon ring).
#include "common.cuh"
#include "integrate.cuh"
#include "ks_metric.cuh"
#include "metric_plugin.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

struct IntegrateArgs {
    const double* s0;      // (npx, 8)
    long npx;
    int N;                 // iteration cap (rows of the reference's scan)
    StepRule rule;
    double* final_state;   // (npx, 8) or null
    int32_t* nsteps;       // (npx,) or null
    double* r_last;        // (npx,) or null : radius_cal(S[argmax(dt) - 1]) with the reference's negative wrap
    double* S;             // dump: (nrows, npx, 8) or null
    double* dt;            // dump: (nrows, npx)
    long nrows;
    unsigned int* queue;   // zero-initialised ray counter
    unsigned long long* total_steps;  // optional global sum of accepted steps
    // paged dump (single pass, ragged): page p = [PAGE_ROWS][8] states followed by [PAGE_ROWS] dts
    double* pages;
    int* page_next;        // (max_pages,) next page of the same ray or -1
    int* page_first;       // (npx,) first page of each ray
    unsigned int* page_counter;
    unsigned int max_pages;
    int* overflow;         // set to 1 when the page pool is exhausted
};

constexpr int PAGE_ROWS = 32;
constexpr int PAGE_DOUBLES = PAGE_ROWS * 9;
constexpr unsigned PAGE_SLAB = 64;       // pages a warp takes from the global pool per atomic
enum { MODE_FINAL = 0, MODE_PADDED = 1, MODE_PAGED = 2 };

template <class Metric, int MODE>
__global__ void __launch_bounds__(128, 4) integrate_kernel(const Metric g, const IntegrateArgs A)
{
    constexpr bool DUMP = (MODE == MODE_PADDED);
    unsigned slab_next = 0, slab_end = 0;      // warp-uniform page slab (MODE_PAGED)
    int page = -1;
    const unsigned lane = threadIdx.x & 31u;
    long ray = -1;
    bool drained = false;           // queue exhausted (warp-uniform)
    double s[8];
    double dt = 0.0, r_cur = 0.0, r_prev = 0.0;
    typename Metric::Cache cache, cache_new;
    double best_dt = 0.0, r_before_best = 0.0;
    int it = 0, best_idx = -1;
    unsigned long long my_steps = 0;

    for (;;) {
        // ---- refill idle lanes from the queue ----
        unsigned idle = __ballot_sync(FULL_MASK, ray < 0);
        if (idle) {
            if (!drained) {
                int cnt = __popc(idle);
                unsigned base = 0;
                int leader = __ffs(idle) - 1;
                if ((int)lane == leader) base = atomicAdd(A.queue, (unsigned)cnt);
                base = __shfl_sync(FULL_MASK, base, leader);
                if ((long)base + cnt >= A.npx) drained = true;
                if (ray < 0) {
                    long idx = (long)base + __popc(idle & ((1u << lane) - 1u));
                    if (idx < A.npx) {
                        ray = idx;
                        const double4* p = reinterpret_cast<const double4*>(A.s0 + idx * 8);
                        double4 lo = p[0], hi = p[1];
                        s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
                        s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
                        r_cur = g.radius(s, cache);
                        dt = A.rule(r_cur);
                        r_prev = r_cur;
                        it = 0; best_idx = -1; best_dt = -1.0e300; r_before_best = r_cur;
                        page = -1;
                    }
                }
            }
            if (__ballot_sync(FULL_MASK, ray >= 0) == 0) break;
        }
        const bool act = ray >= 0;

        // ---- paged dump: lanes starting a new page take one from the warp's slab ----
        if (MODE == MODE_PAGED) {
            bool need = act && ((it & (PAGE_ROWS - 1)) == 0);
            unsigned nm = __ballot_sync(FULL_MASK, need);
            if (nm) {
                unsigned cnt = (unsigned)__popc(nm);
                if (slab_next + cnt > slab_end) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(A.page_counter, PAGE_SLAB);
                    slab_next = __shfl_sync(FULL_MASK, base, 0);
                    slab_end = slab_next + PAGE_SLAB;
                }
                if (need) {
                    unsigned np = slab_next + (unsigned)__popc(nm & ((1u << lane) - 1u));
                    if (np < A.max_pages) {
                        if (it == 0) A.page_first[ray] = (int)np;
                        else if (page >= 0) A.page_next[page] = (int)np;
                        A.page_next[np] = -1;
                        page = (int)np;
                    } else {
                        *A.overflow = 1;
                        if (it == 0) A.page_first[ray] = -1;
                        page = -1;
                    }
                }
                slab_next += cnt;
            }
        }
        if (!act) continue;

        // ---- one iteration of geodesic_step ----
        double cand[8];
        double r_new = 0.0, dtn = 0.0;
        if (dt != 0.0) {
            rk4_step(g, s, dt, cand, &cache);
            r_new = g.radius(cand, cache_new);
            dtn = A.rule(r_new);
        }
        bool frozen = (dt == 0.0) || (dtn == 0.0);
        if (DUMP) {
            if (it < A.nrows) {
                double4* p = reinterpret_cast<double4*>(A.S + ((long)it * A.npx + ray) * 8);
                p[0] = make_double4(s[0], s[1], s[2], s[3]);
                p[1] = make_double4(s[4], s[5], s[6], s[7]);
                A.dt[(long)it * A.npx + ray] = frozen ? 0.0 : dt;
            }
        }
        if (MODE == MODE_PAGED) {
            if (page >= 0) {
                double* pg = A.pages + (long)page * PAGE_DOUBLES;
                int rr = it & (PAGE_ROWS - 1);
                double4* p = reinterpret_cast<double4*>(pg + rr * 8);
                p[0] = make_double4(s[0], s[1], s[2], s[3]);
                p[1] = make_double4(s[4], s[5], s[6], s[7]);
                pg[PAGE_ROWS * 8 + rr] = frozen ? 0.0 : dt;
            }
        }
        bool done = frozen;
        // geodesics.py:373: argmax(dt) is the first zero row (= it) unless some step size was positive (a ray
        // that jumped inside the horizon steps with dt > 0); the classifier row is argmax - 1, and -1 wraps to
        // the last row, a copy of the frozen state.
        double rl = (best_dt > 0.0) ? ((best_idx >= 1) ? r_before_best : r_cur) : ((it >= 1) ? r_prev : r_cur);
        if (!frozen) {
            if (dt > best_dt) { best_dt = dt; best_idx = it; r_before_best = r_prev; }
            r_prev = r_cur; r_cur = r_new;
            cache = cache_new;
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = cand[i];
            dt = dtn;
            it++;
            if (it == A.N) {                        // never froze: argmax over the negative dts (:373)
                done = true;
                rl = (best_idx >= 1) ? r_before_best : r_prev;
            }
        }
        if (done) {
            if (A.final_state) {
                double4* p = reinterpret_cast<double4*>(A.final_state + ray * 8);
                p[0] = make_double4(s[0], s[1], s[2], s[3]);
                p[1] = make_double4(s[4], s[5], s[6], s[7]);
            }
            if (A.nsteps) A.nsteps[ray] = it;
            if (A.r_last) A.r_last[ray] = rl;
            my_steps += (unsigned long long)it;
            ray = -1;
        }
    }
    if (A.total_steps) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_steps += __shfl_xor_sync(FULL_MASK, my_steps, o);
        if (lane == 0 && my_steps) atomicAdd(A.total_steps, my_steps);
    }
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

template <class Metric>
static int launch_integrate(const Metric& g, const IntegrateArgs& A, cudaStream_t stream)
{
    int per_sm = 0;
    auto kern = A.pages ? integrate_kernel<Metric, MODE_PAGED>
                        : (A.S ? integrate_kernel<Metric, MODE_PADDED> : integrate_kernel<Metric, MODE_FINAL>);
    MK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0));
    if (per_sm < 1) per_sm = 1;
    long warps_needed = (A.npx + 31) / 32;
    long blocks = (long)sm_count() * per_sm;
    long need = (warps_needed + 3) / 4;
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 128, 0, stream>>>(g, A);
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
    if (npx == 0 || N == 0) return 0;
    IntegrateArgs A;
    A.s0 = s0; A.npx = npx; A.N = (int)N;
    A.rule.div = div; A.rule.inv_div = 1.0 / div; A.rule.tol = tol;
    A.final_state = final_state; A.nsteps = nsteps; A.r_last = r_last;
    A.S = S; A.dt = dt; A.nrows = S ? nrows : 0;
    A.queue = queue_counter(stream, 0);
    if (!A.queue) return 1;
    A.total_steps = total_steps;
    A.pages = nullptr; A.page_next = nullptr; A.page_first = nullptr; A.page_counter = nullptr;
    A.max_pages = 0; A.overflow = nullptr;
    return dispatch_integrate(metric_id, bhspin, A, stream);
}

static int dispatch_integrate(int metric_id, double bhspin, IntegrateArgs& A, cudaStream_t stream)
{
    int rc;
    if (metric_id == MK_METRIC_KERR_SCHILD) {
        KerrSchild g; g.a = bhspin; g.aa = bhspin * bhspin; g.rH = 1.0 + sqrt(1.0 - bhspin * bhspin);
        A.rule.rH = g.rH;
        rc = launch_integrate(g, A, stream);
    } else if (metric_id == MK_METRIC_KERR_SCHILD_DUAL) {
        DualMetric<KerrSchildFn> g; g.fn.a = bhspin; g.rH = 1.0 + sqrt(1.0 - bhspin * bhspin);
        A.rule.rH = g.rH;
        rc = launch_integrate(g, A, stream);
    } else {
        set_error("unknown metric id %d (runtime-registered metrics go through mk_integrate_plugin)", metric_id);
        return 2;
    }
    return rc;
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
extern "C" int mk_page_rows(void) { return PAGE_ROWS; }

extern "C" int mk_integrate_paged(int metric_id, double bhspin, long N, long npx, const double* s0, double div,
                                  double tol, double* final_state, int32_t* nsteps, double* r_last,
                                  double* pages, int32_t* page_next, int32_t* page_first,
                                  unsigned int* page_counter, long max_pages, int32_t* overflow,
                                  unsigned long long* total_steps, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MK_REQUIRE(npx >= 0 && N >= 0 && N < (1L << 31) - 2, "npx / N out of range");
    MK_REQUIRE(pages && page_next && page_first && page_counter && overflow, "null page-store pointer");
    MK_REQUIRE(max_pages > 0 && max_pages < (1L << 31), "max_pages out of range");
    MK_REQUIRE(npx == 0 || s0 != nullptr, "s0 is null");
    MK_REQUIRE(div != 0.0, "div must be non-zero");
    if (npx == 0 || N == 0) return 0;
    IntegrateArgs A;
    A.s0 = s0; A.npx = npx; A.N = (int)N;
    A.rule.div = div; A.rule.inv_div = 1.0 / div; A.rule.tol = tol;
    A.final_state = final_state; A.nsteps = nsteps; A.r_last = r_last;
    A.S = nullptr; A.dt = nullptr; A.nrows = 0;
    A.queue = queue_counter(stream, 0);
    if (!A.queue) return 1;
    A.total_steps = total_steps;
    A.pages = pages; A.page_next = page_next; A.page_first = page_first; A.page_counter = page_counter;
    A.max_pages = (unsigned)max_pages; A.overflow = overflow;
    return dispatch_integrate(metric_id, bhspin, A, stream);
}

namespace mk {
// paged store -> the reference's padded layout for a selection of rays: S (nrows, nsel, 8), dt (nrows, nsel).
// One thread per selected ray walks its page chain; rows past the ray's last stored row repeat it with dt = 0.
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
    int page = page_first[ray];
    double4 lo = make_double4(0, 0, 0, 0), hi = lo;
    for (long row = 0; row < nrows; row++) {
        double d = 0.0;
        if (row < have && page >= 0) {
            int rr = (int)(row & (PAGE_ROWS - 1));
            const double* pg = pages + (long)page * PAGE_DOUBLES;
            const double4* p = reinterpret_cast<const double4*>(pg + rr * 8);
            lo = p[0]; hi = p[1];
            d = pg[PAGE_ROWS * 8 + rr];
            if (rr == PAGE_ROWS - 1) page = page_next[page];
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
