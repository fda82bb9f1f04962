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
//            shared memory (straight-line: 13 stores and one mbarrier arrive per lane);
//   warps 1-3 (consumers) take slots round-robin, do block lookup + trilinear gather + fluid frame + j_nu, alpha_nu for
//            their lane's record, then fold it IN ORDER into the lane's (I, T): they read (I, T) as the previous slot
//            left it (waiting for that slot's mbarrier), apply the two FMAs of the transfer update and leave the result
//            in their own slot; only those two FMAs are serial across the consumers;
//   warp 0   reads the last slot's I at the end of the patch and stores the pixel.
// Every number is produced by the same device functions with the same operands in the same order as in render_body, so
// the pixels are bit-identical to the fused kernel's (tests/test_fluid_gpu.py::test_long_patch_pipeline_is_bit_identical).
// With 16 warps resident per SM the producer advances at its share of the FP64 pipe (455 instead of 795 FP64 instructions
// per step on its chain), and once the GPU drains at 0.73 us per step (0.645 for the geodesic alone).
//
// Built-in Kerr-Schild spacetime, 1 to 8 observing frequencies (the consumers evaluate all of them for a sample and
// fold each into its own (I_f, T_f)); registered spacetimes go through render_body.  Replaces nothing new in the reference: same contract as render_kernel.cuh
// (/root/reference/mahakala/images.py:56-144).
#pragma once
#include "render_kernel.cuh"

namespace mk {

constexpr int PIPE_RING = 8;            // slots per CTA
constexpr int PIPE_CONSUMERS = 3;       // warps 1..3
constexpr int PIPE_IN_FIELDS = 13;      // s[8], f, l1, l2, l3, weight
// fields per slot: the 13 inputs, overwritten after use by the lane's (I, T) of every frequency (2 NF values)
constexpr int pipe_fields(int nf) { return 2 * nf > PIPE_IN_FIELDS ? 2 * nf : PIPE_IN_FIELDS; }
constexpr int PIPE_THREADS = 32 * (1 + PIPE_CONSUMERS);      // threads per patch group
// Patch groups per CTA.  1: a 128-thread CTA that shares its SM with whatever else is resident (e.g. three CTAs of the
// bulk launch, whose warps then compete with the producer for FP64 issue slots).  4: a 512-thread CTA = the whole
// register file of an SM, nothing else can be resident; the warps are arranged so that every SM sub-partition holds
// ONE producer and three consumers of the other groups (warp w sits on sub-partition w % 4: group = w / 4, the
// producer of group g is warp 5 g).
constexpr int PIPE_GROUPS_EXCLUSIVE = 4;

// Per-group shared memory, addressed with 32-bit shared-window addresses (st.shared / ld.shared / mbarrier on
// shared::cta): generic pointers to dynamic shared memory make the compiler rebuild the window base from SR_CgaCtaId at
// every use, which a lone latency-bound warp pays for in full.
//   rec   [PIPE_RING][FIELDS][32] f64        a lane touches only its own column; fields 0 .. 2 NF - 1 are overwritten with
//                                            the lane's (I_f, T_f) after this slot
//   full  [PIPE_RING] mbarrier (32 arrivals) producer lanes -> consumer
//   done  [PIPE_RING] mbarrier (32 arrivals) consumer lanes -> next consumer (fold) and producer (slot reuse)
//   mask  [PIPE_RING] u32                    lanes that carry a record
//   type  [PIPE_RING] i32                    0 = samples, 2 = samples, first slot of a patch, 1 = exit
template <int NF>
struct PipeLayout {
    static constexpr unsigned FIELDS = pipe_fields(NF);
    static constexpr unsigned OFF_FULL = PIPE_RING * FIELDS * 32 * 8;
    static constexpr unsigned OFF_DONE = OFF_FULL + PIPE_RING * 8;
    static constexpr unsigned OFF_MASK = OFF_DONE + PIPE_RING * 8;
    static constexpr unsigned OFF_TYPE = OFF_MASK + PIPE_RING * 4;
    static constexpr unsigned GROUP_BYTES = OFF_TYPE + PIPE_RING * 4;
};

__device__ __forceinline__ void pipe_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pipe_mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");      // release.cta: the lane's stores first
}
__device__ __forceinline__ void pipe_mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok = 0;
    while (!ok)     // try_wait suspends the warp in hardware for a bounded time: waiting warps do not spin on issue slots
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned pipe_mbar_test(unsigned bar, unsigned parity)       // non-blocking
{
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void pipe_sts(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ double pipe_lds(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void pipe_sts32(unsigned addr, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned pipe_lds32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// conservative "could this point be sampled?" (a superset of the exact membership test the consumer applies)
__device__ __forceinline__ bool pipe_maybe_inside(const SnapshotView& sn, const double s[8])
{
    // branch-free: every divergent region between the first RK4 stage and the rest of the step costs the lone producer
    // its reconvergence
    return (sn.source != 0) | ((sn.bbox_lo[0] <= s[1]) & (s[1] <= sn.bbox_hi[0]) & (sn.bbox_lo[1] <= s[2]) &
                               (s[2] <= sn.bbox_hi[1]) & (sn.bbox_lo[2] <= s[3]) & (s[3] <= sn.bbox_hi[2]));
}

template <int NF, int KIND, int GROUPS>
__device__ __forceinline__ void render_pipeline_body(const KerrSchild& G, const RenderArgs& A)
{
    typedef PipeLayout<NF> LY;
    constexpr unsigned PIPE_FIELDS = LY::FIELDS, PIPE_OFF_FULL = LY::OFF_FULL, PIPE_OFF_DONE = LY::OFF_DONE,
                       PIPE_OFF_MASK = LY::OFF_MASK, PIPE_OFF_TYPE = LY::OFF_TYPE, PIPE_GROUP_BYTES = LY::GROUP_BYTES;
    extern __shared__ __align__(16) unsigned char mk_pipe_smem[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned group = 0, role = warp;                // role 0 = producer, 1..3 = consumers
    // (128-thread CTAs: rotating the producer warp by the CTA's wave index, (blockIdx.x / SMs) & 3, so that the four
    // resident CTAs of an SM would not all have their producer in warp 0, was tried and measured SLOWER -- 1.24 against
    // 1.13 us per step at four patches per SM; the block scheduler does not deal CTAs out in that order.)
    if (GROUPS > 1) {
        const unsigned q = warp >> 2, r = warp & 3u;
        group = q;
        role = (r == q) ? 0u : 1u + (r > q ? r - 1u : r);
    }
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(mk_pipe_smem);
    if (threadIdx.x == 0) {
        for (int g = 0; g < GROUPS; g++)
            for (int k = 0; k < 2 * PIPE_RING; k++) pipe_mbar_init(smem0 + g * PIPE_GROUP_BYTES + PIPE_OFF_FULL + 8 * k, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // exclusive CTAs may run with fewer than four working groups: the CTA still owns the SM's register file (nothing
    // else becomes resident), and the producers share the FP64 pipe with fewer neighbours
    if (GROUPS > 1 && (int)group >= A.pipe_groups) return;
    const unsigned sh = smem0 + group * PIPE_GROUP_BYTES;       // this group's block
    const unsigned col = sh + lane * 8;                         // + (slot * PIPE_FIELDS + field) * 256: this lane's cell
    auto rec = [&](unsigned k, unsigned q) { return col + (k * PIPE_FIELDS + q) * 256u; };

    if (role != 0) {
        // ------------------------------------------------ consumers ------------------------------------------------
        unsigned long long my_samples = 0;
        for (unsigned n = role - 1;; n += PIPE_CONSUMERS) {
            const unsigned k = n % PIPE_RING;
            pipe_mbar_wait(sh + PIPE_OFF_FULL + 8 * k, (n / PIPE_RING) & 1u);
            const unsigned ty = pipe_lds32(sh + PIPE_OFF_TYPE + 4 * k);
            if (ty == 1u) break;
            const unsigned m = pipe_lds32(sh + PIPE_OFF_MASK + 4 * k);
            bool inside = false;
            double e_out[NF], a_out[NF], w = 0.0;
#pragma unroll
            for (int fq = 0; fq < NF; fq++) { e_out[fq] = 0.0; a_out[fq] = 0.0; }
            if ((m >> lane) & 1u) {
                double s[8];
#pragma unroll
                for (int q = 0; q < 8; q++) s[q] = pipe_lds(rec(k, q));
                const double f = pipe_lds(rec(k, 8));
                const double l[4] = {1.0, pipe_lds(rec(k, 9)), pipe_lds(rec(k, 10)), pipe_lds(rec(k, 11))};
                w = pipe_lds(rec(k, 12));
                double prims[8];
#ifdef MK_PIPE_NOCONSUME       // experiment: consumers that do nothing (pixels are wrong)
                if (false) {
#else
                if (interp_prims_kind<KIND>(A.sn, s, prims)) {
#endif
                    inside = true;
                    emission_fast<NF>(A.P, A.C, f, l, s, prims, A.nu_obs, A.inv_nu_obs,
                                      [&](int fq, double e, double a) { e_out[fq] = e; a_out[fq] = a; });
                }
            }
            // in-order fold: (I, T) after the previous slot of this patch (written by whichever consumer had it) -> after
            // this one.  Only these two FMAs are serial across the consumers; the sample above is not.
            const unsigned kp = (n - 1u) % PIPE_RING;
            if (ty != 2u) pipe_mbar_wait(sh + PIPE_OFF_DONE + 8 * kp, ((n - 1u) / PIPE_RING) & 1u);
#pragma unroll
            for (int fq = 0; fq < NF; fq++) {
                double I = 0.0, T = 1.0;
                if (ty != 2u) {
                    I = pipe_lds(rec(kp, 2 * fq));
                    T = pipe_lds(rec(kp, 2 * fq + 1));
                }
                if (inside) {
                    const double Tf = T;             // the update of render_body, operand for operand
                    I = fma(Tf, w * e_out[fq], I);
                    T = Tf * fma(-w, a_out[fq], 1.0);
                }
                pipe_sts(rec(k, 2 * fq), I);
                pipe_sts(rec(k, 2 * fq + 1), T);
            }
            pipe_mbar_arrive(sh + PIPE_OFF_DONE + 8 * k);
            const unsigned im = __ballot_sync(FULL_MASK, inside);
            if (lane == 0) my_samples += (unsigned long long)__popc(im);
        }
        if (lane == 0 && A.total_samples && my_samples) atomicAdd(A.total_samples, my_samples);
        return;
    }

    // -------------------------------------------------- producer --------------------------------------------------
    unsigned push_ptr = 0, done_ptr = 0;     // slots pushed / slots known to be finished (both warp-uniform)
    unsigned long long my_steps = 0;
    // Slot m may overwrite slot m - RING only when slot m - RING + 1 is finished too (its consumer reads the (I, T) that
    // slot m - RING left behind): the producer keeps push_ptr + 2 <= done_ptr + RING before every push.

    for (;;) {
        unsigned pq = 0;
        if (lane == 0) pq = atomicAdd_system(A.queue, 1u);
        pq = __shfl_sync(FULL_MASK, pq, 0);
        long patch = A.patch_begin + (long)pq * A.patch_stride;
        if (patch >= A.patch_end) break;
        if (A.patch_order) patch = A.patch_order[patch];

        long ray;
        double s[8] = {0.0, 100.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0};       // lanes without a ray step on this (discarded)
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
        const unsigned patch_first = push_ptr;
        int it = 0;
        double dt = 0.0;
        KerrSchild::Cache cache;
        const double dt_first = A.rule(G.radius(s, cache));
        if (active) dt = dt_first;
        if (dt == 0.0) active = false;
        double wdt = 0.0;
        bool pending = false;
        while (__any_sync(FULL_MASK, active)) {
            double a1[4];
            KerrSchild::MetricFunctions mf = {0.0, 0.0, 0.0, 0.0};
            // room for one push: when the ring is one slot short, the oldest slot's barrier is TESTED before the first
            // RK4 stage and its answer used after it, so the round trip to the barrier hides behind ~90 FP64 operations
            // instead of stalling the chain (an iteration pushes at most one slot, so one retirement keeps up)
            const bool need_room = push_ptr + 2u > done_ptr + (unsigned)PIPE_RING;
            const unsigned old_bar = sh + PIPE_OFF_DONE + 8 * (done_ptr % PIPE_RING), old_par = (done_ptr / PIPE_RING) & 1u;
            unsigned room_ok = 1u;
            if (need_room) room_ok = pipe_mbar_test(old_bar, old_par);
            // No divergent regions in this loop: every lane executes the step, finished lanes on whatever their registers
            // hold (no memory access depends on it, nothing of theirs is kept) -- a lone producer pays for every
            // reconvergence point between the first RK4 stage, the push and the rest of the step (measured: 0.85 ->
            // 0.80 us per step for the branch-free bounding-box test alone).
            G.accel(s, s + 4, a1, &cache, &mf);
            if (need_room & (room_ok == 0u)) pipe_mbar_wait(old_bar, old_par);       // rare: the consumers are behind
            done_ptr += need_room ? 1u : 0u;
            // hand the state to the consumers (render_body samples it here)
#ifdef MK_PIPE_NOPUSH          // experiment: the producer alone (no slot is ever pushed; pixels are wrong)
            const bool want = false;
#else
            const bool want = active & pending & pipe_maybe_inside(A.sn, s);
#endif
            const unsigned pm = __ballot_sync(FULL_MASK, want);
            if (pm) {
                const unsigned k = push_ptr % PIPE_RING;
                // straight-line: every lane stores its column (the mask says which ones count) and arrives itself, the
                // two control words are written by all lanes with the same value
#pragma unroll
                for (int q = 0; q < 8; q++) pipe_sts(rec(k, q), s[q]);
                // metric functions of the sample: whatever render_body<NF> feeds its fluid-frame algebra -- the first
                // RK4 stage's (f, l) in the stage-1-first loop, the point cache's in the plain loop (bit-identity)
                if constexpr (NF > MK_RENDER_PIPE_MAX) G.fl(s, cache, mf.f, mf.l1, mf.l2, mf.l3);
                pipe_sts(rec(k, 8), mf.f); pipe_sts(rec(k, 9), mf.l1); pipe_sts(rec(k, 10), mf.l2);
                pipe_sts(rec(k, 11), mf.l3); pipe_sts(rec(k, 12), wdt);
                pipe_sts32(sh + PIPE_OFF_MASK + 4 * k, pm);
                pipe_sts32(sh + PIPE_OFF_TYPE + 4 * k, (push_ptr == patch_first) ? 2u : 0u);
                pipe_mbar_arrive(sh + PIPE_OFF_FULL + 8 * k);
                push_ptr++;
            }
            rk4_rest(G, s, a1, dt, s);
            const double dtn = A.rule(G.radius(s, cache));
            // geodesics.py:264-267: a step whose end point fails the rule is rejected and the ray is frozen
            const bool moved = active & (dtn != 0.0);
            wdt = moved ? -dt * A.P.L_unit : wdt;       // -dt[i-1] * L_unit  (> 0): weight of the sample at the new state
            dt = moved ? dtn : dt;
            it += moved ? 1 : 0;
            pending = pending | moved;
            active = moved & (it != A.N);               // row N is not part of the reference's scan output
        }
        while (done_ptr != push_ptr) {
            pipe_mbar_wait(sh + PIPE_OFF_DONE + 8 * (done_ptr % PIPE_RING), (done_ptr / PIPE_RING) & 1u);
            done_ptr++;
        }
        if (valid) {
#pragma unroll
            for (int fq = 0; fq < NF; fq++)
                A.image[(long)fq * A.npx + ray] = (push_ptr != patch_first) ? pipe_lds(rec((push_ptr - 1u) % PIPE_RING, 2 * fq)) : 0.0;
            if (A.nsteps) A.nsteps[ray] = it;
            my_steps += (unsigned long long)it;
        }
    }
    // release the consumers: one exit slot for each of them (consecutive slot numbers cover all residues)
    for (int j = 0; j < PIPE_CONSUMERS; j++) {
        const unsigned k = push_ptr % PIPE_RING;
        pipe_sts32(sh + PIPE_OFF_MASK + 4 * k, 0u);
        pipe_sts32(sh + PIPE_OFF_TYPE + 4 * k, 1u);
        pipe_mbar_arrive(sh + PIPE_OFF_FULL + 8 * k);
        push_ptr++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_steps += __shfl_xor_sync(FULL_MASK, my_steps, o);
    if (lane == 0 && A.total_steps && my_steps) atomicAdd(A.total_steps, my_steps);
}

}  // namespace mk
