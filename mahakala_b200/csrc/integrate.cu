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
};

template <class Metric, bool DUMP>
__global__ void __launch_bounds__(128, 4) integrate_kernel(const Metric g, const IntegrateArgs A)
{
    const unsigned lane = threadIdx.x & 31u;
    long ray = -1;
    bool drained = false;           // queue exhausted (warp-uniform)
    double s[8];
    double dt = 0.0, r_cur = 0.0, r_prev = 0.0;
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
                        r_cur = g.radius(s);
                        dt = A.rule(r_cur);
                        r_prev = r_cur;
                        it = 0; best_idx = -1; best_dt = -1.0e300; r_before_best = r_cur;
                    }
                }
            }
            if (__ballot_sync(FULL_MASK, ray >= 0) == 0) break;
        }
        if (ray < 0) continue;

        // ---- one iteration of geodesic_step ----
        double cand[8];
        double r_new = 0.0, dtn = 0.0;
        if (dt != 0.0) {
            rk4_step(g, s, dt, cand);
            r_new = g.radius(cand);
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
        bool done = frozen;
        double rl = (it >= 1) ? r_prev : r_cur;     // first zero row = it ; classifier row = it - 1 (wraps)
        if (!frozen) {
            if (dt > best_dt) { best_dt = dt; best_idx = it; r_before_best = r_prev; }
            r_prev = r_cur; r_cur = r_new;
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
    auto kern = A.S ? integrate_kernel<Metric, true> : integrate_kernel<Metric, false>;
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
