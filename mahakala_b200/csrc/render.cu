// Fused render kernel: camera ray -> RK4 geodesic -> snapshot sample -> j_nu, alpha_nu -> intensity.
//
// Replaces the whole chunk loop of /root/reference/mahakala/images.py:56-144 (initialize_geodesics_at_camera,
// geodesic_integrator, get_fluid_scalars_from_geodesics, rlow_rhigh_model, synchrotron_coefficients, sigma
// cut, solve_specific_intensity).  Nothing of shape (nrows, npx, .) is ever materialised: each lane keeps
// its ray's state and the (I, T) accumulators of every observing frequency in registers.
//
// Transfer order.  The reference accumulates back to front (transfer.py:106-119):
//     for i = n .. 1:  I <- I (1 - a_i) + s_i ,  s_i = -dt_{i-1} L j_i ,  a_i = -dt_{i-1} L alpha_i
// which is the linear recurrence  I = sum_i s_i prod_{m<i} (1 - a_m).  The kernel marches camera -> hole,
// so it evaluates the same sum front to back:  I += T s_i ; T *= (1 - a_i).  (Identical in exact
// arithmetic; rounding differs at the 1e-16 level per term, tests bound the per-pixel difference.)
//
// Scheduling.  Persistent CTAs; every warp pulls 32-ray patches (4 x 8 pixels of the grid camera, so that
// the lanes of a warp traverse the same snapshot cells at the same time) from a global atomic queue.  The
// queue counter and the image may live in a peer GPU's memory: several GPUs then share ONE dynamic tile
// queue over NVLink and write finished pixels straight into the gathering rank's image.
#include <cstdlib>
#include "common.cuh"
#include "camera.cuh"
#include "integrate.cuh"
#include "ks_metric.cuh"
#include "snapshot.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

constexpr int PATCH_X = 4, PATCH_Y = 8;      // pixels per warp patch: 4 (ix) x 8 (iy)

struct RenderArgs {
    KerrSchild g;
    CameraGeom cam;
    double fov_lo, step;
    long res, patches_y;
    const double* s0;          // explicit rays (npx, 8) or null for the grid camera
    long npx;
    int N;
    StepRule rule;
    SnapshotView sn;
    EmissionParams P;
    EmissionConsts C;
    double nu_obs[8], inv_nu_obs[8];
    double* image;             // (NF, npx)
    int32_t* nsteps;
    unsigned long long* total_steps;
    unsigned long long* total_samples;
    unsigned int* queue;
    long patch_begin, patch_end, patch_stride;
    const int* patch_order;    // optional permutation of the patch indices (scheduling order)
};

// Resident CTAs per SM.  Measured on B200 (scripts/dev/render_variants.py, cfg4, 1 / 2 / 8 frequencies): 4 CTAs of
// 128 threads at 128 registers 26.8 / 28.8 / 39.5 ms; 3 CTAs at 168 registers 27.2 / 29.8 / 39.5 ms; 13-15 warps per
// SM with 136-152 registers (one- or two-warp CTAs) 27.7-28.1 / 30.2-30.5 / 39.7-41.6 ms.  The plateau is flat
// (+-2 %): the kernel is bound by dependent FP64 latency plus FP64 issue, and occupancy trades against spills.
#ifndef MK_RENDER_SPLIT
#define MK_RENDER_SPLIT 4
#endif
#ifndef MK_RENDER_LO
#define MK_RENDER_LO 4
#endif
#ifndef MK_RENDER_THREADS
#define MK_RENDER_THREADS 128
#endif
#ifndef MK_RENDER_HI
#define MK_RENDER_HI 4
#endif
#ifndef MK_RENDER_PIPE_MAX
#define MK_RENDER_PIPE_MAX 0
#endif
// Experiment knob (off: 9 > max NF): from this many frequencies on, the (I, T) accumulators of a lane live in shared
// memory ([2 NF][threads], conflict free) instead of registers.  Measured on B200 (cfg4, 8 frequencies): 40.1 ms
// against 39.3 ms with register accumulators -- the ~200 B of spills of the 8-frequency kernel come from the RK4 /
// emission temporaries under the 128-register cap, not from the accumulators.
#ifndef MK_RENDER_SMEM_MIN
#define MK_RENDER_SMEM_MIN 9
#endif
#ifdef MK_RENDER_MAXREG        // experiment: cap registers directly (any warp count per SM with small CTAs)
#define MK_RENDER_BOUNDS __maxnreg__(MK_RENDER_MAXREG)
#else
#define MK_RENDER_BOUNDS __launch_bounds__(MK_RENDER_THREADS, (NF >= MK_RENDER_SPLIT) ? MK_RENDER_HI : MK_RENDER_LO)
#endif
template <int NF, int KIND>
__global__ void MK_RENDER_BOUNDS render_kernel(const RenderArgs A)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long my_steps = 0, my_samples = 0;

    for (;;) {
        // ---- next patch ----
        unsigned pq = 0;
        // system scope: the counter may live in a peer GPU's memory (one queue shared by all GPUs of the node), and
        // only system-scope atomics are guaranteed atomic across devices; one atomic per 32-ray patch either way
        if (lane == 0) pq = atomicAdd_system(A.queue, 1u);
        pq = __shfl_sync(FULL_MASK, pq, 0);
        long patch = A.patch_begin + (long)pq * A.patch_stride;
        if (patch >= A.patch_end) break;
        if (A.patch_order) patch = A.patch_order[patch];

        long ray;
        double s[8];
        bool active;
        if (A.s0) {
            ray = patch * 32 + lane;
            active = ray < A.npx;
            if (active) {
                const double4* p = reinterpret_cast<const double4*>(A.s0 + ray * 8);
                double4 lo = p[0], hi = p[1];
                s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
                s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
            }
        } else {
            long px = patch / A.patches_y, py = patch - px * A.patches_y;
            long ix = px * PATCH_X + (lane >> 3), iy = py * PATCH_Y + (lane & 7u);
            active = ix < A.res && iy < A.res;
            ray = ix * A.res + iy;
            if (active) {
                double x[4], v[4];
                camera_point(A.cam, pixel_centre(A.fov_lo, A.step, ix), pixel_centre(A.fov_lo, A.step, iy), x, v);
                nullify_state(A.g, x, v, s);
            }
        }
        const bool valid = active;
        constexpr bool SMEM_ACC = (NF >= MK_RENDER_SMEM_MIN);
        __shared__ double sacc[SMEM_ACC ? 2 * NF * MK_RENDER_THREADS : 1];
        double Ireg[SMEM_ACC ? 1 : NF], Treg[SMEM_ACC ? 1 : NF];
        auto I = [&](int f) -> double& { return SMEM_ACC ? sacc[(2 * f) * MK_RENDER_THREADS + threadIdx.x] : Ireg[SMEM_ACC ? 0 : f]; };
        auto T = [&](int f) -> double& { return SMEM_ACC ? sacc[(2 * f + 1) * MK_RENDER_THREADS + threadIdx.x] : Treg[SMEM_ACC ? 0 : f]; };
#pragma unroll
        for (int f = 0; f < NF; f++) { I(f) = 0.0; T(f) = 1.0; }
        int it = 0;
        double dt = 0.0;
        KerrSchild::Cache cache;
        if (active) dt = A.rule(A.g.radius(s, cache));
        if (dt == 0.0) active = false;          // never moves: n = 0, no row pair contributes

        // Two loop shapes (compile-time, MK_RENDER_PIPE_MAX = largest NF that uses the first):
        //  * stage-1-first: the first RK4 stage of the step that LEAVES state s is evaluated before s is sampled, and
        //    its metric functions (f, l) feed the fluid-frame algebra of the sample, so the sample needs no metric
        //    evaluation of its own (17 FP64 operations and the dependency on the point cache);
        //  * plain: "step, then sample the new state" with f, l from the point cache.
        // (The ping-pong register scheme of integrate_kernel.cuh, which removes the s = cand copies, was tried here
        // too in round 1: it duplicates the whole sample + emission + RK4 body, and the kernel got 25 % SLOWER -- 34.2 vs
        // 27.4 ms on cfg4 -- at any register budget: the doubled code no longer fits the instruction cache.)
        if constexpr (NF <= MK_RENDER_PIPE_MAX) {
            double wdt = 0.0;
            bool pending = false;
            while (__any_sync(FULL_MASK, active)) {
                if (active) {
                    double a1[4];
                    KerrSchild::MetricFunctions mf;
                    A.g.accel(s, s + 4, a1, &cache, &mf);
                    if (pending) {
                        double prims[8];
                        if (interp_prims_kind<KIND>(A.sn, s, prims)) {
                            my_samples++;
                            const double l[4] = {1.0, mf.l1, mf.l2, mf.l3};
                            emission_fast<NF>(A.P, A.C, mf.f, l, s, prims, A.nu_obs, A.inv_nu_obs,
                                              [&](int fq, double e, double a) {
                                                  const double Tf = T(fq);
                                                  I(fq) = fma(Tf, wdt * e, I(fq));
                                                  T(fq) = Tf * fma(-wdt, a, 1.0);
                                              });
                        }
                    }
                    // in place: when the step is rejected the ray retires and its old state (sampled above) is not
                    // needed any more
                    rk4_rest(A.g, s, a1, dt, s);
                    const double dtn = A.rule(A.g.radius(s, cache));
                    if (dtn == 0.0) {
                        active = false;             // step rejected; ray frozen (geodesics.py:264-267)
                    } else {
                        wdt = -dt * A.P.L_unit;     // -dt[i-1] * L_unit  (> 0): weight of the sample at the new state
                        dt = dtn;
                        it++;
                        pending = true;
                        if (it == A.N) active = false;      // row N is not part of the reference's scan output
                    }
                }
            }
        } else {
            while (__any_sync(FULL_MASK, active)) {
                if (active) {
                    rk4_step(A.g, s, dt, s, &cache);            // in place: a rejected step retires the ray
                    double dtn = A.rule(A.g.radius(s, cache));
                    if (dtn == 0.0) {
                        active = false;             // step rejected; ray frozen (geodesics.py:264-267)
                    } else {
                        const double wdt = -dt * A.P.L_unit;     // -dt[i-1] * L_unit  (> 0)
                        dt = dtn;
                        it++;
                        if (it == A.N) {
                            active = false;         // row N is not part of the reference's scan output
                        } else {
                            double prims[8];
                            if (interp_prims_kind<KIND>(A.sn, s, prims)) {
                                my_samples++;
                                double f, l[4];
                                l[0] = 1.0;
                                A.g.fl(s, cache, f, l[1], l[2], l[3]);
                                // each frequency is folded into (I, T) as soon as its coefficients exist
                                // (em = ab = 0 leaves them unchanged)
                                emission_fast<NF>(A.P, A.C, f, l, s, prims, A.nu_obs, A.inv_nu_obs,
                                                  [&](int fq, double e, double a) {
                                                      const double Tf = T(fq);
                                                      I(fq) = fma(Tf, wdt * e, I(fq));
                                                      T(fq) = Tf * fma(-wdt, a, 1.0);
                                                  });
                            }
                        }
                    }
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int fq = 0; fq < NF; fq++) A.image[(long)fq * A.npx + ray] = I(fq);
            if (A.nsteps) A.nsteps[ray] = it;
            my_steps += (unsigned long long)it;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_steps += __shfl_xor_sync(FULL_MASK, my_steps, o);
        my_samples += __shfl_xor_sync(FULL_MASK, my_samples, o);
    }
    if (lane == 0) {
        if (A.total_steps && my_steps) atomicAdd(A.total_steps, my_steps);
        if (A.total_samples && my_samples) atomicAdd(A.total_samples, my_samples);
    }
}

#ifdef MK_EXPERIMENTS
// EXPERIMENT (compiled only with -DMK_EXPERIMENTS, e.g. scripts/build_variant.sh "-DMK_EXPERIMENTS"; not part
// of the product library): lane-level refill for the fused kernel.  Idle lanes take single pixels from a
// pixel-granular queue (same centre-out patch order) as soon as at least `thr` lanes of the warp are idle.
// Measured on B200 (cfg4, scripts/dev/refill_probe.py): thr = 1 / 8 / 16 / 24 / 32 -> 54.1 / 48.6 / 40.5 / 34.8 /
// 28.2 ms against 27.7 ms for whole patches, i.e. incoherent warps cost far more than idle tail lanes.
__global__ void __launch_bounds__(128, 3) render_refill_kernel(const RenderArgs A, int thr)
{
    const unsigned lane = threadIdx.x & 31u;
    long ray = -1;
    bool drained = false;
    double s[8], I = 0.0, T = 1.0, dt = 0.0, wdt = 0.0;
    int it = 0;
    bool pending = false;
    KerrSchild::Cache cache, cache_new;
    const long total = (A.patch_end - A.patch_begin) * 32;
    for (;;) {
        unsigned idle = __ballot_sync(FULL_MASK, ray < 0);
        if (idle) {
            if (!drained && (__popc(idle) >= thr || idle == FULL_MASK)) {
                int cnt = __popc(idle);
                unsigned base = 0;
                int leader = __ffs(idle) - 1;
                if ((int)lane == leader) base = atomicAdd(A.queue, (unsigned)cnt);
                base = __shfl_sync(FULL_MASK, base, leader);
                if ((long)base + cnt >= total) drained = true;
                if (ray < 0) {
                    long q = (long)base + __popc(idle & ((1u << lane) - 1u));
                    if (q < total) {
                        long patch = A.patch_begin + (q >> 5);
                        if (A.patch_order) patch = A.patch_order[patch];
                        unsigned k = (unsigned)(q & 31);
                        long px = patch / A.patches_y, py = patch - px * A.patches_y;
                        long ix = px * PATCH_X + (k >> 3), iy = py * PATCH_Y + (k & 7u);
                        if (ix < A.res && iy < A.res) {
                            ray = ix * A.res + iy;
                            double x[4], v[4];
                            camera_point(A.cam, pixel_centre(A.fov_lo, A.step, ix), pixel_centre(A.fov_lo, A.step, iy), x, v);
                            nullify_state(A.g, x, v, s);
                            dt = A.rule(A.g.radius(s, cache));
                            I = 0.0; T = 1.0; it = 0; pending = false;
                            if (dt == 0.0) { A.image[ray] = 0.0; ray = -1; }
                        }
                    }
                }
            }
            if (__ballot_sync(FULL_MASK, ray >= 0) == 0) {
                if (drained) break;
                continue;
            }
        }
        if (ray < 0) continue;
        double cand[8], prims[8], dtn;
        if (pending && interp_prims_kind<KIND>(A.sn, s, prims)) {
            double f, l[4], em[1], ab[1];
            l[0] = 1.0;
            A.g.fl(s, cache, f, l[1], l[2], l[3]);
            emission_fast<1>(A.P, A.C, f, l, s, prims, A.nu_obs, A.inv_nu_obs,
                             [&](int, double e, double a) { em[0] = e; ab[0] = a; });
            rk4_step(A.g, s, dt, cand, &cache);
            dtn = A.rule(A.g.radius(cand, cache_new));
            I = fma(T, wdt * em[0], I);
            T = T * fma(-wdt, ab[0], 1.0);
        } else {
            rk4_step(A.g, s, dt, cand, &cache);
            dtn = A.rule(A.g.radius(cand, cache_new));
        }
        bool done = (dtn == 0.0);
        if (!done) {
            wdt = -dt * A.P.L_unit;
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = cand[i];
            cache = cache_new;
            dt = dtn;
            it++;
            pending = true;
            if (it == A.N) done = true;
        }
        if (done) {
            A.image[ray] = I;
            ray = -1;
        }
    }
}

#endif  // MK_EXPERIMENTS

template <int NF, int KIND>
static int launch_render_kind(const RenderArgs& A, long npatches, cudaStream_t stream);

template <int NF>
static int launch_render(const RenderArgs& A, long npatches, cudaStream_t stream)
{
    switch (snapshot_kind(A.sn)) {
        case SNAP_F64_GRID_POW2: return launch_render_kind<NF, SNAP_F64_GRID_POW2>(A, npatches, stream);
        case SNAP_F32_GRID_POW2: return launch_render_kind<NF, SNAP_F32_GRID_POW2>(A, npatches, stream);
        default: return launch_render_kind<NF, SNAP_GENERIC>(A, npatches, stream);
    }
}

template <int NF, int KIND>
static int launch_render_kind(const RenderArgs& A, long npatches, cudaStream_t stream)
{
#ifdef MK_EXPERIMENTS
    if (NF == 1 && !A.s0) {
        static int thr = -2;
        if (thr == -2) { const char* e = getenv("MK_RENDER_REFILL_THR"); thr = e ? atoi(e) : -1; }
        if (thr >= 1) {
            long blocks = (long)sm_count() * 3;
            render_refill_kernel<<<(unsigned)blocks, 128, 0, stream>>>(A, thr);
            MK_CUDA_CHECK(cudaGetLastError());
            return 0;
        }
    }
#endif
    int per_sm = 0;
    MK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel<NF, KIND>, MK_RENDER_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    long blocks = (long)sm_count() * per_sm;
    const long warps = MK_RENDER_THREADS / 32;
    long need = (npatches + warps - 1) / warps;
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    render_kernel<NF, KIND><<<(unsigned)blocks, MK_RENDER_THREADS, 0, stream>>>(A);
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

extern "C" int mk_render(double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
                         double fov_upper, long res, const double* s0, long npx, long N, double div,
                         double tol, const mk_snapshot* snap, const mk_emission_params* params, int nfreq,
                         const double* nu_obs, double* image, int32_t* nsteps,
                         unsigned long long* total_steps, unsigned long long* total_samples,
                         unsigned int* queue, long patch_begin, long patch_end, long patch_stride,
                         const int* patch_order, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MK_REQUIRE(snap && params && nu_obs && image, "null pointer");
    MK_REQUIRE(nfreq >= 1 && nfreq <= 8, "nfreq must be in 1..8");
    MK_REQUIRE(N >= 0 && N < (1L << 31) - 2, "N out of range");
    MK_REQUIRE(div != 0.0, "div must be non-zero");
    if (!s0) { MK_REQUIRE(res >= 0 && npx == res * res, "npx must equal res*res for the grid camera"); }
    if (npx == 0) return 0;
    RenderArgs A;
    A.g.set_spin(bhspin);
    A.cam.ci = cos_i; A.cam.si = sin_i; A.cam.d = distance;
    A.fov_lo = fov_lower;
    A.step = res > 0 ? (fov_upper - fov_lower) / (double)(2 * res) : 0.0;
    A.res = res;
    A.patches_y = (res + PATCH_Y - 1) / PATCH_Y;
    A.s0 = s0; A.npx = npx; A.N = (int)N;
    A.rule.div = div; A.rule.inv_div = 1.0 / div; A.rule.tol = tol; A.rule.rH = A.g.rH;
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
    if (A.patch_begin >= A.patch_end) return 0;
    A.queue = queue ? queue : queue_counter(stream, 1);
    if (!A.queue) return 1;
    long span = (A.patch_end - A.patch_begin + A.patch_stride - 1) / A.patch_stride;
    if (nfreq == 1) return launch_render<1>(A, span, stream);
    if (nfreq == 2) return launch_render<2>(A, span, stream);
    if (nfreq == 3) return launch_render<3>(A, span, stream);
    if (nfreq == 4) return launch_render<4>(A, span, stream);
    if (nfreq == 5) return launch_render<5>(A, span, stream);
    if (nfreq == 6) return launch_render<6>(A, span, stream);
    if (nfreq == 7) return launch_render<7>(A, span, stream);
    return launch_render<8>(A, span, stream);
}
