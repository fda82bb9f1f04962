// Warp-specialised variant of the fused render kernel for the LONG patches of a frame (photon-ring rays).
//
// Why.  A ray is a chain of dependent RK4 steps (up to 3765 at BASELINE cfg2/cfg4); in render_body the snapshot sample
// and the emission chain of every step sit on that same chain.  Measured on B200 with nothing else on the GPU
// (scripts/dev/lone_warp_probe.py): 0.64 us per step for the geodesic alone, 1.95 us per step with sample + emission --
// the sample is two thirds of the latency of a long ray, and the longest patch alone (7.3 ms) is what stops ONE small
// frame from strong-scaling over 8 GPUs (1024^2 frame: 23 ms on one GPU, 7.7-8.7 ms on eight).
//
// What.  The sample never feeds back into the geodesic, so it does not have to be on its critical path.  One CTA = one
// patch at a time = four warps:
//   warp 0   (producer)  integrates the 32 rays exactly like render_body's stage-1-first loop and, instead of sampling,
//            pushes {state, f, l, weight} of every lane that may be inside the snapshot into a ring of PIPE_RING slots in
//            shared memory (one elected mbarrier arrive per slot);
//   warps 1-3 (consumers) take slots round-robin, do block lookup + trilinear gather + fluid frame + j_nu, alpha_nu for
//            their lane's record and write (j, alpha) back into the slot;
//   warp 0   folds finished slots IN ORDER into its lane's (I, T) accumulators (two FMAs per sample) when the ring slot is
//            needed again, and at the end of the patch.
// Every number is produced by the same device functions with the same operands in the same order as in render_body, so
// the pixels are bit-identical to the fused kernel's (tests/test_fluid_gpu.py::test_long_patch_pipeline_is_bit_identical).
// With 16 warps resident per SM the producer advances at its share of the FP64 pipe (455 instead of 795 FP64 instructions
// per step on its chain), and once the GPU drains at 0.64 us per step.
//
// Built-in Kerr-Schild spacetime, one observing frequency (the latency-critical single-frame case); everything else
// goes through render_body.  Replaces nothing new in the reference: same contract as render_kernel.cuh
// (/root/reference/mahakala/images.py:56-144).
#pragma once
#include "render_kernel.cuh"

namespace mk {

constexpr int PIPE_RING = 8;            // slots per CTA
constexpr int PIPE_CONSUMERS = 3;       // warps 1..3
constexpr int PIPE_FIELDS = 13;         // s[8], f, l1, l2, l3, weight
constexpr int PIPE_THREADS = 32 * (1 + PIPE_CONSUMERS);

struct PipeShared {
    double rec[PIPE_RING][PIPE_FIELDS][32];         // [slot][field][lane]: a lane touches only its own column
    unsigned long long full[PIPE_RING];             // producer -> consumer
    unsigned long long done[PIPE_RING];             // consumer -> producer
    unsigned mask[PIPE_RING];                       // lanes that carry a record
    unsigned inside[PIPE_RING];                     // lanes whose record was inside the snapshot (contribute)
    int type[PIPE_RING];                            // 0 = samples, 1 = exit
};

__device__ __forceinline__ unsigned pipe_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pipe_mbar_init(unsigned long long* b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(pipe_smem_addr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void pipe_mbar_arrive(unsigned long long* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(pipe_smem_addr(b)) : "memory");
}
__device__ __forceinline__ void pipe_mbar_wait(unsigned long long* b, unsigned parity)
{
    unsigned ok = 0;
    while (!ok)     // try_wait suspends the warp in hardware for a bounded time: waiting warps do not spin on issue slots
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(pipe_smem_addr(b)), "r"(parity) : "memory");
}

// conservative "could this point be sampled?" (a superset of the exact membership test the consumer applies)
__device__ __forceinline__ bool pipe_maybe_inside(const SnapshotView& sn, const double s[8])
{
    if (sn.source != 0) return true;
    return (sn.bbox_lo[0] <= s[1]) & (s[1] <= sn.bbox_hi[0]) & (sn.bbox_lo[1] <= s[2]) & (s[2] <= sn.bbox_hi[1]) &
           (sn.bbox_lo[2] <= s[3]) & (s[3] <= sn.bbox_hi[2]);
}

template <int KIND>
__device__ __forceinline__ void render_pipeline_body(const KerrSchild& G, const RenderArgs& A)
{
    __shared__ PipeShared sh;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int k = 0; k < PIPE_RING; k++) { pipe_mbar_init(&sh.full[k], 1); pipe_mbar_init(&sh.done[k], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp != 0) {
        // ------------------------------------------------ consumers ------------------------------------------------
        unsigned long long my_samples = 0;
        for (unsigned n = warp - 1;; n += PIPE_CONSUMERS) {
            const unsigned k = n % PIPE_RING;
            pipe_mbar_wait(&sh.full[k], (n / PIPE_RING) & 1u);
            if (sh.type[k] != 0) break;
            const unsigned m = sh.mask[k];
            bool inside = false;
            if ((m >> lane) & 1u) {
                double s[8];
#pragma unroll
                for (int q = 0; q < 8; q++) s[q] = sh.rec[k][q][lane];
                const double f = sh.rec[k][8][lane];
                const double l[4] = {1.0, sh.rec[k][9][lane], sh.rec[k][10][lane], sh.rec[k][11][lane]};
                double prims[8];
                if (interp_prims_kind<KIND>(A.sn, s, prims)) {
                    inside = true;
                    double e_out = 0.0, a_out = 0.0;
                    emission_fast<1>(A.P, A.C, f, l, s, prims, A.nu_obs, A.inv_nu_obs,
                                     [&](int, double e, double a) { e_out = e; a_out = a; });
                    sh.rec[k][0][lane] = e_out;
                    sh.rec[k][1][lane] = a_out;
                }
            }
            const unsigned im = __ballot_sync(FULL_MASK, inside);
            if (lane == 0) { sh.inside[k] = im; my_samples += (unsigned long long)__popc(im); }
            __syncwarp();
            if (lane == 0) pipe_mbar_arrive(&sh.done[k]);
        }
        if (lane == 0 && A.total_samples && my_samples) atomicAdd(A.total_samples, my_samples);
        return;
    }

    // -------------------------------------------------- producer --------------------------------------------------
    unsigned push_ptr = 0, fold_ptr = 0;
    double I = 0.0, T = 1.0;
    unsigned long long my_steps = 0;
    auto fold_one = [&]() {
        const unsigned k = fold_ptr % PIPE_RING;
        pipe_mbar_wait(&sh.done[k], (fold_ptr / PIPE_RING) & 1u);
        if ((sh.inside[k] >> lane) & 1u) {
            const double e = sh.rec[k][0][lane], a = sh.rec[k][1][lane], w = sh.rec[k][12][lane];
            const double Tf = T;                     // the update of render_body, operand for operand
            I = fma(Tf, w * e, I);
            T = Tf * fma(-w, a, 1.0);
        }
        fold_ptr++;
    };

    for (;;) {
        unsigned pq = 0;
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
        I = 0.0; T = 1.0;
        int it = 0;
        double dt = 0.0;
        KerrSchild::Cache cache;
        if (active) dt = A.rule(G.radius(s, cache));
        if (dt == 0.0) active = false;
        double wdt = 0.0;
        bool pending = false;
        while (__any_sync(FULL_MASK, active)) {
            double a1[4];
            KerrSchild::MetricFunctions mf;
            if (active) G.accel(s, s + 4, a1, &cache, &mf);
            // hand the state to the consumers (render_body samples it here)
            const bool want = active && pending && pipe_maybe_inside(A.sn, s);
            const unsigned pm = __ballot_sync(FULL_MASK, want);
            if (pm) {
                while (push_ptr - fold_ptr >= (unsigned)PIPE_RING) fold_one();
                const unsigned k = push_ptr % PIPE_RING;
                if (want) {
#pragma unroll
                    for (int q = 0; q < 8; q++) sh.rec[k][q][lane] = s[q];
                    sh.rec[k][8][lane] = mf.f; sh.rec[k][9][lane] = mf.l1; sh.rec[k][10][lane] = mf.l2;
                    sh.rec[k][11][lane] = mf.l3; sh.rec[k][12][lane] = wdt;
                }
                if (lane == 0) { sh.mask[k] = pm; sh.type[k] = 0; }
                __syncwarp();
                if (lane == 0) pipe_mbar_arrive(&sh.full[k]);
                push_ptr++;
            }
            if (active) {
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
        while (fold_ptr != push_ptr) fold_one();
        if (valid) {
            A.image[ray] = I;
            if (A.nsteps) A.nsteps[ray] = it;
            my_steps += (unsigned long long)it;
        }
    }
    // release the consumers: one exit slot for each of them (consecutive slot numbers cover all residues)
    for (int j = 0; j < PIPE_CONSUMERS; j++) {
        const unsigned k = push_ptr % PIPE_RING;
        if (lane == 0) { sh.mask[k] = 0u; sh.type[k] = 1; }
        __syncwarp();
        if (lane == 0) pipe_mbar_arrive(&sh.full[k]);
        push_ptr++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_steps += __shfl_xor_sync(FULL_MASK, my_steps, o);
    if (lane == 0 && A.total_steps && my_steps) atomicAdd(A.total_steps, my_steps);
}

}  // namespace mk
