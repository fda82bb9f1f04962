// Device-side body of the fused render kernel, generic in the spacetime (shared by the built-in kernels of render.cu
// and by run-time compiled metric plugins, plugin_tu.cuh).  NVRTC-safe: no host code, no standard headers.
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
//
// Metric concept used here (KerrSchild, DualMetric<Fn>):  accel / radius / Cache as in integrate_kernel.cuh, plus
//     render_nullify(g, x, v, s)                       initial_condition with that spacetime (geodesics.py:219-230)
//     render_frame(g, s, cache, prims) -> FrameScalars fluid-frame scalars k.u, k.b, b.b (athenak.py:760-786)
#pragma once
#include "camera.cuh"
#include "integrate.cuh"
#include "ks_metric.cuh"
#include "metric_plugin.cuh"
#include "sample.cuh"

namespace mk {

constexpr int PATCH_X = 4, PATCH_Y = 8;      // pixels per warp patch: 4 (ix) x 8 (iy)

struct RenderArgs {
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
    int* nsteps;               // (npx,) int32 or null
    unsigned long long* total_steps;
    unsigned long long* total_samples;
    unsigned int* queue;
    long patch_begin, patch_end, patch_stride;
    const int* patch_order;    // optional permutation of the patch indices (scheduling order)
    int pipe_groups;           // long-patch kernel, exclusive CTAs: patch groups that work (1..4), the others exit at once
};


// ---- spacetime adaptors ----
__device__ __forceinline__ void render_nullify(const KerrSchild& g, const double x[4], const double v[4], double s[8])
{
    nullify_state(g, x, v, s);
}
__device__ __forceinline__ FrameScalars render_frame(const KerrSchild& g, const double s[8], const KerrSchild::Cache& cache,
                                                     const double prims[8])
{
    double f, l[4];
    l[0] = 1.0;
    g.fl(s, cache, f, l[1], l[2], l[3]);
    return frame_kerr_schild(f, l, s, prims);
}
template <class Fn>
__device__ __forceinline__ void render_nullify(const DualMetric<Fn>& g, const double x[4], const double v[4], double s[8])
{
    double gm[4][4];
    g.fn(x, gm);
    nullify_with_metric(gm, x, v, s);
}
template <class Fn>
__device__ __forceinline__ FrameScalars render_frame(const DualMetric<Fn>& g, const double s[8],
                                                     const typename DualMetric<Fn>::Cache&, const double prims[8])
{
    double gc[4][4], gi[4][4];
    g.metric_cov_con(s, gc, gi);
    return frame_generic(gc, gi, s, prims);
}
template <class Metric> struct has_stage1_metric_functions { static constexpr bool value = false; };
template <> struct has_stage1_metric_functions<KerrSchild> { static constexpr bool value = true; };

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
// Largest NF that uses the stage-1-first loop.  Measured on B200 (cfg4, f64 cells, same box): one frequency 22.96-23.13
// ms against 24.06 ms for the plain loop; two and more frequencies within noise of each other (25.6 / 27.1 / 31.6 vs
// 25.6 / 26.6 / 31.4 ms for 2 / 4 / 8 frequencies), so only the single-frequency kernel takes it.
#ifndef MK_RENDER_PIPE_MAX
#define MK_RENDER_PIPE_MAX 1
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
template <class Metric, int NF, int KIND>
__device__ __forceinline__ void render_body(const Metric& G, const RenderArgs& A)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long my_steps = 0, my_samples = 0;
#ifdef MK_RENDER_SMEM_STAGE
    // experiment: per-warp brick of staged cells + mbarrier in dynamic shared memory (f64 grid snapshots only)
    extern __shared__ __align__(128) unsigned char mk_dyn_smem[];
    CellStage stage;
    stage.brick = reinterpret_cast<double*>(mk_dyn_smem) + (threadIdx.x >> 5) * 216;
    stage.bar = reinterpret_cast<unsigned long long*>(mk_dyn_smem + (MK_RENDER_THREADS / 32) * 1728) + (threadIdx.x >> 5);
    if (KIND == SNAP_F64_GRID_POW2) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(stage.bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
#define MK_INTERP(sn, s, prims) ((KIND == SNAP_F64_GRID_POW2) ? interp_prims_staged(sn, s, prims, stage) : interp_prims_kind<KIND>(sn, s, prims))
#else
#define MK_INTERP(sn, s, prims) interp_prims_kind<KIND>(sn, s, prims)
#endif

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
                render_nullify(G, x, v, s);
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
        typename Metric::Cache cache;
        if (active) dt = A.rule(G.radius(s, cache));
        if (dt == 0.0) active = false;          // never moves: n = 0, no row pair contributes

        // Two loop shapes (compile-time, MK_RENDER_PIPE_MAX = largest NF that uses the first):
        //  * stage-1-first: the first RK4 stage of the step that LEAVES state s is evaluated before s is sampled, and
        //    its metric functions (f, l) feed the fluid-frame algebra of the sample, so the sample needs no metric
        //    evaluation of its own (17 FP64 operations and the dependency on the point cache);
        //  * plain: "step, then sample the new state" with f, l from the point cache.
        // (The ping-pong register scheme of integrate_kernel.cuh, which removes the s = cand copies, was tried here
        // too in round 1: it duplicates the whole sample + emission + RK4 body, and the kernel got 25 % SLOWER -- 34.2 vs
        // 27.4 ms on cfg4 -- at any register budget: the doubled code no longer fits the instruction cache.)
        if constexpr (NF <= MK_RENDER_PIPE_MAX && has_stage1_metric_functions<Metric>::value) {
            double wdt = 0.0;
            bool pending = false;
            while (__any_sync(FULL_MASK, active)) {
                if (active) {
                    double a1[4];
                    KerrSchild::MetricFunctions mf;
#if defined(MK_RENDER_PREFETCH)
                    // experiment: cell addresses first, a prefetch of the four rows, then the first RK4 stage, then the
                    // gather -- the stage's ~90 FP64 operations cover the L2 latency of the rows
                    CellRef cref;
                    bool located = false;
                    if (KIND == SNAP_F64_GRID_POW2 && pending) {
                        located = locate_cells_f64(A.sn, s, cref);
                        if (located) prefetch_cells_f64(A.sn, cref);
                    }
#endif
                    G.accel(s, s + 4, a1, &cache, &mf);
                    if (pending) {
                        double prims[8];
#if defined(MK_RENDER_PREFETCH)
                        bool inside = located;
                        if (KIND == SNAP_F64_GRID_POW2) { if (located) gather_cells_f64(A.sn, cref, prims); }
                        else inside = MK_INTERP(A.sn, s, prims);
                        if (inside) {
#else
                        if (MK_INTERP(A.sn, s, prims)) {
#endif
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
                    rk4_rest(G, s, a1, dt, s);
                    const double dtn = A.rule(G.radius(s, cache));
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
                    rk4_step(G, s, dt, s, &cache);            // in place: a rejected step retires the ray
                    double dtn = A.rule(G.radius(s, cache));
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
                            if (MK_INTERP(A.sn, s, prims)) {
                                my_samples++;
                                // each frequency is folded into (I, T) as soon as its coefficients exist
                                // (em = ab = 0 leaves them unchanged)
                                emission_from_frame<NF>(A.P, A.C, render_frame(G, s, cache, prims), prims,
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

}  // namespace mk
