// Device-side body of the persistent integrate kernel (shared by the built-in kernels of integrate.cu and
// by run-time compiled metric plugins, plugin_tu.cuh).  NVRTC-safe: no host code, no standard headers.
//
// Replaces /root/reference/mahakala/geodesics.py:233-281 (geodesic_integrator) and the last-point
// rule of :370-378.  One ray per lane, state in registers.  The kernel is persistent: each warp pulls
// rays from a global queue and REFILLS lanes whose ray has frozen (ballot + one atomic per refill), so
// warps stay full although step counts vary ~5x across the image (photon ring).
#pragma once
#include "integrate.cuh"

namespace mk {

// 256-bit global store (STG.E.256 on sm_100): one instruction per 32 B instead of two 128-bit stores
__device__ __forceinline__ void store_256(double* p, double a, double b, double c, double d)
{
#if !defined(__CUDACC_RTC__) && defined(MK_DUMP_STREAMING)      // experiment: evict-first streaming stores for the dump
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#elif !defined(__CUDACC_RTC__)
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else   // NVRTC 12.9's embedded ptxas rejects 256-bit vector accesses: two 128-bit stores for run-time plugins
    reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
#endif
}

struct IntegrateArgs {
    const double* s0;      // (npx, 8)
    long npx;
    int N;                 // iteration cap (rows of the reference's scan)
    StepRule rule;
    double* final_state;   // (npx, 8) or null
    int* nsteps;              // (npx,) or null
    double* r_last;        // (npx,) or null : radius_cal(S[argmax(dt) - 1]) with the reference's negative wrap
    double* S;             // dump: (nrows, npx, 8) or null
    double* dt;            // dump: (nrows, npx)
    long nrows;
    unsigned int* queue;   // zero-initialised ray counter; may live in a peer GPU's memory (one queue for all GPUs)
    int chunk_div;         // guided self-scheduling: a warp asks for (rays left) / chunk_div positions, 1..QUEUE_CHUNK
    unsigned chunk_mul;    // floor(2^32 / chunk_div): the division above as one multiply-high
    const int* ray_order;  // optional: queue position q -> ray index (scheduling order; results stay indexed by ray)
    int page_id_offset;    // added to the page numbers stored in page_first (rank * pool size in a multi-GPU job)
    unsigned long long* total_steps;  // optional global sum of accepted steps
    // paged dump (single pass, ragged).  Each WARP appends to its own log: in every loop iteration the active
    // lanes write their row into the same slot, so one slot is 32 x 64 B = 2 KB of contiguous state (fully
    // coalesced 256-bit stores) plus 32 contiguous step sizes.  Page p = [PAGE_SLOTS][32][8] states followed
    // by [PAGE_SLOTS][32] dts; a warp's pages are chained through page_next.  A ray is the column `lane` of
    // consecutive slots starting at (page, slot) = page_first[ray] (lanes are refilled, so a column holds
    // several rays one after another).
    double* pages;
    int* page_next;        // (max_pages,) next page of the same warp or -1
    int* page_first;       // (npx, 2): first page of the ray, slot * 32 + lane of its row 0
    unsigned int* page_counter;
    unsigned int max_pages;
    int* overflow;         // set to 1 when the page pool is exhausted
};

constexpr int PAGE_SLOTS = 16;
// Paged dump by direct 256-bit global stores (0, the product path) or through shared memory + TMA bulk stores (1, an
// EXPERIMENT kept behind this switch; always 0 for run-time compiled metric plugins).  Measured on B200, cfg2, same
// box, 4 CTAs/SM: direct stores 16.03 ms, TMA staging 17.28 ms (final-state mode without any dump: 14.9 ms).  The
// staging costs more than it hides: the 64 B rows of 32 lanes conflict 4-way in shared memory (the layout in
// shared memory has to be the layout in the page, a bulk copy is linear), plus two warp barriers and a proxy fence
// per iteration, while the direct 256-bit stores were never the bottleneck of the write path -- the HBM write
// stream itself is (2.5 TB/s of the 3.9 TB/s a write-only kernel reaches on this part).
#ifndef MK_DUMP_TMA
#define MK_DUMP_TMA 0
#endif
#ifdef __CUDACC_RTC__
#undef MK_DUMP_TMA
#define MK_DUMP_TMA 0
#endif
constexpr int DUMP_SLOT_BYTES = 32 * 72;            // 32 states of 64 B followed by 32 step sizes
constexpr int DUMP_SMEM_PER_WARP = 2 * DUMP_SLOT_BYTES;
// Rays are taken from the queue in chunks of up to QUEUE_CHUNK positions per warp, and the NEXT chunk is requested
// while the current one is being consumed: the atomic's round trip (a few microseconds when the counter sits in a
// peer GPU's memory across NVLink) overlaps the RK4 steps in between instead of stalling the warp at every refill.
// Chunks shrink as the queue empties (guided self-scheduling: rays left / (4 x warps of all participants)), down to
// single rays, so that no warp sits on a private stock of rays while lanes elsewhere run dry.
constexpr int QUEUE_CHUNK = 32;
constexpr int PAGE_DOUBLES = PAGE_SLOTS * 32 * 9;
enum { MODE_FINAL = 0, MODE_PADDED = 1, MODE_PAGED = 2 };

// Per-lane ray bookkeeping that survives across steps (everything except the state vector and its point cache,
// which ping-pong between two register sets, see integrate_body).
struct LaneRay {
    long ray;
    double dt, r_cur, r_prev, best_dt, r_before_best;
    int it, best_idx;
};

// SHARED = false: the queue is this GPU's own counter -- a lane refill costs one device-scope atomic for exactly the
// idle lanes (ray-granular, nothing is held back).  SHARED = true: the counter may sit in a peer GPU's memory --
// chunked, prefetched system-scope atomics with guided chunk sizes and an optional scheduling order (see above).
// (Measured on B200, cfg2, one GPU: the chunked scheme costs ~1.5 % when the queue is local, hence two variants.)
// One ray from s0 to its end, no queue: the per-lane logic of integrate_body (same calls in the same order, so the
// same numbers) for callers that iterate on the outcome of single rays (shadow bisection).  Returns the classifier
// radius radius_cal(S[argmax(dt) - 1]) of geodesics.py:370-378; s0 is overwritten with the final state.
template <class Metric>
MK_HD double integrate_one(const Metric& g, const StepRule& rule, double (&s)[8], int N, int& nsteps)
{
    typename Metric::Cache c, cn;
    double sn[8];
    double r_cur = g.radius(s, c), r_prev = r_cur, best_dt = -1.0e300, r_before_best = r_cur;
    double dt = rule(r_cur);
    int it = 0, best_idx = -1;
    bool capped = false;
    for (;;) {
        double r_new = 0.0, dtn = 0.0;
        if (dt != 0.0) {
            rk4_step(g, s, dt, sn, &c);
            r_new = g.radius(sn, cn);
            dtn = rule(r_new);
        }
        if ((dt == 0.0) || (dtn == 0.0)) break;
        if (dt > best_dt) { best_dt = dt; best_idx = it; r_before_best = r_prev; }
        r_prev = r_cur; r_cur = r_new; dt = dtn; it++;
#pragma unroll
        for (int m = 0; m < 8; m++) s[m] = sn[m];
        c = cn;
        if (it == N) { capped = true; break; }
    }
    nsteps = it;
    if (capped) return (best_idx >= 1) ? r_before_best : r_prev;
    return (best_dt > 0.0) ? ((best_idx >= 1) ? r_before_best : r_cur) : ((it >= 1) ? r_prev : r_cur);
}

template <class Metric, int MODE, bool SHARED = false>
__device__ __forceinline__ void integrate_body(const Metric& g, const IntegrateArgs& A)
{
    constexpr bool DUMP = (MODE == MODE_PADDED);
    int wpage = -1, wslot = 0, my_slot = 0;    // warp-uniform log position (MODE_PAGED)
    int dump_count = 0;                        // slots handed to the TMA engine so far (warp-uniform)
    const unsigned lane = threadIdx.x & 31u;
    bool drained = false;           // queue exhausted (warp-uniform)
    long cur = 0, cur_end = 0;      // the warp's private range of queue positions (warp-uniform)
    unsigned nxt = 0;               // base of the prefetched chunk (meaningful in lane 0 while have_next)
    int nxt_len = 0;                // its length (warp-uniform)
    long last_base = 0;             // base of the latest chunk received: how far the queue has advanced
    bool have_next = false;
    LaneRay L;
    L.ray = -1; L.dt = 0.0; L.r_cur = 0.0; L.r_prev = 0.0; L.best_dt = 0.0; L.r_before_best = 0.0; L.it = 0; L.best_idx = -1;
    unsigned long long my_steps = 0;

    auto request_chunk = [&]() {
        unsigned want = __umulhi((unsigned)(A.npx - last_base), A.chunk_mul);     // rays left / chunk_div (npx < 2^31)
        nxt_len = want < 1u ? 1 : (want > (unsigned)QUEUE_CHUNK ? QUEUE_CHUNK : (int)want);
        if (lane == 0) nxt = atomicAdd_system(A.queue, (unsigned)nxt_len);
        have_next = true;
    };

    // One loop iteration of the reference's scan for every lane of the warp: refill idle lanes into (s, cache),
    // take one step of the active lanes from (s, cache) into (sn, cn).  Returns false when the warp is done.
    // The caller alternates the two register sets, so an accepted step needs no register-to-register copy of
    // the 8-vector and its cache (the loop-carried moves were ~8 % of the issued instructions).
    auto iteration = [&](double (&s)[8], typename Metric::Cache& cache, double (&sn)[8],
                         typename Metric::Cache& cn) -> bool {
        unsigned idle = __ballot_sync(FULL_MASK, L.ray < 0);
        if (!SHARED && idle) {
            // ---- refill idle lanes from the local queue ----
            if (!drained) {
                int cnt = __popc(idle);
                unsigned base = 0;
                int leader = __ffs(idle) - 1;
                if ((int)lane == leader) base = atomicAdd(A.queue, (unsigned)cnt);
                base = __shfl_sync(FULL_MASK, base, leader);
                if ((long)base + cnt >= A.npx) drained = true;
                if (L.ray < 0) {
                    long idx = (long)base + __popc(idle & ((1u << lane) - 1u));
                    if (idx < A.npx) {
                        L.ray = idx;
                        const double4* p = reinterpret_cast<const double4*>(A.s0 + idx * 8);
                        double4 lo = p[0], hi = p[1];
                        s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
                        s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
                        L.r_cur = g.radius(s, cache);
                        L.dt = A.rule(L.r_cur);
                        L.r_prev = L.r_cur;
                        L.it = 0; L.best_idx = -1; L.best_dt = -1.0e300; L.r_before_best = L.r_cur;
                    }
                }
            }
            if (__ballot_sync(FULL_MASK, L.ray >= 0) == 0) return false;
        }
        if (SHARED && idle) {
            // ---- refill idle lanes from the warp's private range, topping it up from the shared queue ----
            int need = __popc(idle);
            const int my_rank = __popc(idle & ((1u << lane) - 1u));     // position among the idle lanes
            int given = 0;
            while (need > 0 && !drained) {
                if (cur == cur_end) {
                    if (!have_next) request_chunk();
                    const unsigned b = __shfl_sync(FULL_MASK, nxt, 0);   // waits for the atomic only if still in flight
                    have_next = false;
                    if ((long)b >= A.npx) { drained = true; break; }
                    cur = last_base = (long)b;
                    cur_end = ((long)b + nxt_len < A.npx) ? (long)b + nxt_len : A.npx;
                }
                const int avail = (int)(cur_end - cur);
                const int take = need < avail ? need : avail;
                if (L.ray < 0 && my_rank >= given && my_rank < given + take) {
                    const long q = cur + (my_rank - given);
                    const long idx = A.ray_order ? (long)A.ray_order[q] : q;
                    L.ray = idx;
                    const double4* p = reinterpret_cast<const double4*>(A.s0 + idx * 8);
                    double4 lo = p[0], hi = p[1];
                    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
                    s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
                    L.r_cur = g.radius(s, cache);
                    L.dt = A.rule(L.r_cur);
                    L.r_prev = L.r_cur;
                    L.it = 0; L.best_idx = -1; L.best_dt = -1.0e300; L.r_before_best = L.r_cur;
                }
                cur += take; given += take; need -= take;
            }
            if (!have_next && !drained && cur_end - cur <= 2) request_chunk();
            if (__ballot_sync(FULL_MASK, L.ray >= 0) == 0) return false;
        }
        const bool act = L.ray >= 0;

        // ---- paged dump: the warp claims the next slot of its log (a new page every PAGE_SLOTS iterations) ----
        if (MODE == MODE_PAGED) {
            if (wslot == 0) {
                unsigned np = 0;
                if (lane == 0) np = atomicAdd(A.page_counter, 1u);
                np = __shfl_sync(FULL_MASK, np, 0);
                if (np < A.max_pages) {
                    if (lane == 0) {
                        if (wpage >= 0) A.page_next[wpage] = (int)np;
                        A.page_next[np] = -1;
                    }
                    wpage = (int)np;
                } else {
                    if (lane == 0) *A.overflow = 1;
                    wpage = -1;
                }
            }
            my_slot = wslot;
            wslot = (wslot + 1) & (PAGE_SLOTS - 1);
            if (act && L.it == 0) {
                A.page_first[2 * L.ray] = wpage < 0 ? -1 : wpage + A.page_id_offset;
                A.page_first[2 * L.ray + 1] = my_slot * 32 + (int)lane;
            }
        }
        if (!(MODE == MODE_PAGED && MK_DUMP_TMA) && !act) return true;

        // ---- one iteration of geodesic_step ----
        double r_new = 0.0, dtn = 0.0;
        if (act && L.dt != 0.0) {
            rk4_step(g, s, L.dt, sn, &cache);
            r_new = g.radius(sn, cn);
            dtn = A.rule(r_new);
        }
        bool frozen = (L.dt == 0.0) || (dtn == 0.0);
        if (DUMP) {
            if (L.it < A.nrows) {
                double* p = A.S + ((long)L.it * A.npx + L.ray) * 8;
                store_256(p, s[0], s[1], s[2], s[3]);
                store_256(p + 4, s[4], s[5], s[6], s[7]);
                A.dt[(long)L.it * A.npx + L.ray] = frozen ? 0.0 : L.dt;
            }
        }
        if (MODE == MODE_PAGED && !MK_DUMP_TMA) {
            if (wpage >= 0) {
                double* pg = A.pages + (long)wpage * PAGE_DOUBLES;
                int rr = my_slot * 32 + (int)lane;
                store_256(pg + rr * 8, s[0], s[1], s[2], s[3]);
                store_256(pg + rr * 8 + 4, s[4], s[5], s[6], s[7]);
                pg[PAGE_SLOTS * 32 * 8 + rr] = frozen ? 0.0 : L.dt;
            }
        }
#if MK_DUMP_TMA && !defined(__CUDACC_RTC__)
        if (MODE == MODE_PAGED) {
            // The warp's slot (32 rows of 64 B + 32 step sizes) is assembled in shared memory and leaves through the
            // TMA engine as two bulk copies (2048 B + 256 B, UBLKCP in SASS): the lanes hand their rows to shared
            // memory and go on stepping, they never wait on the HBM write path (2.5 TB/s of the 3.9 TB/s write-only
            // bandwidth) as they do with direct stores.  Two buffers per warp; a buffer is reused once the bulk copy
            // issued from it two iterations ago has read it (wait_group.read 1).  All 32 lanes arrive here (idle lanes
            // skipped the step above), so lane 0 is always the issuer and owns the bulk async-groups.
            if (wpage >= 0) {
                extern __shared__ __align__(128) unsigned char mk_dump_smem[];
                double* sbuf = reinterpret_cast<double*>(mk_dump_smem + ((threadIdx.x >> 5) * 2 + (dump_count & 1)) * DUMP_SLOT_BYTES);
                if (lane == 0 && dump_count >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                if (act) {
                    double2* row = reinterpret_cast<double2*>(sbuf + lane * 8);
                    row[0] = make_double2(s[0], s[1]); row[1] = make_double2(s[2], s[3]);
                    row[2] = make_double2(s[4], s[5]); row[3] = make_double2(s[6], s[7]);
                    sbuf[256 + lane] = frozen ? 0.0 : L.dt;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    double* pg = A.pages + (long)wpage * PAGE_DOUBLES;
                    const unsigned src = (unsigned)__cvta_generic_to_shared(sbuf);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2048;"
                                 :: "l"(pg + my_slot * 256), "r"(src) : "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;"
                                 :: "l"(pg + PAGE_SLOTS * 256 + my_slot * 32), "r"(src + 2048u) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                dump_count++;
            }
            if (!act) return true;
        }
#endif
        bool done = frozen;
        bool capped = false;
        if (!frozen) {
            if (L.dt > L.best_dt) { L.best_dt = L.dt; L.best_idx = L.it; L.r_before_best = L.r_prev; }
            L.r_prev = L.r_cur; L.r_cur = r_new;
            L.dt = dtn;
            L.it++;
            if (L.it == A.N) {                      // never froze
                done = true;
                capped = true;
            }
        }
        if (done) {
            // a frozen ray ends at s (the rejected candidate is discarded); a capped ray at the accepted sn
            if (A.final_state) {
                double4* p = reinterpret_cast<double4*>(A.final_state + L.ray * 8);
                p[0] = capped ? make_double4(sn[0], sn[1], sn[2], sn[3]) : make_double4(s[0], s[1], s[2], s[3]);
                p[1] = capped ? make_double4(sn[4], sn[5], sn[6], sn[7]) : make_double4(s[4], s[5], s[6], s[7]);
            }
            if (A.nsteps) A.nsteps[L.ray] = L.it;
            if (A.r_last) {
                // geodesics.py:373: argmax(dt) is the first zero row (= it) unless some step size was positive (a
                // ray that jumped inside the horizon steps with dt > 0); the classifier row is argmax - 1, and -1
                // wraps to the last row, a copy of the frozen state.  A ray that never froze (capped) has no zero
                // row: argmax over its negative dts.  (For a frozen ray r_prev / r_cur are still those of s.)
                double rl;
                if (capped) rl = (L.best_idx >= 1) ? L.r_before_best : L.r_prev;
                else rl = (L.best_dt > 0.0) ? ((L.best_idx >= 1) ? L.r_before_best : L.r_cur)
                                            : ((L.it >= 1) ? L.r_prev : L.r_cur);
                A.r_last[L.ray] = rl;
            }
            my_steps += (unsigned long long)L.it;
            L.ray = -1;
        }
        return true;
    };

    double sa[8], sb[8];
    typename Metric::Cache ca, cb;
    for (;;) {
        if (!iteration(sa, ca, sb, cb)) break;      // live state: set A -> set B
        if (!iteration(sb, cb, sa, ca)) break;      // live state: set B -> set A
    }
#if MK_DUMP_TMA && !defined(__CUDACC_RTC__)
    if (MODE == MODE_PAGED && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all slots have landed
#endif
    (void)dump_count;
    if (A.total_steps) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_steps += __shfl_xor_sync(FULL_MASK, my_steps, o);
        if (lane == 0 && my_steps) atomicAdd(A.total_steps, my_steps);
    }
}

}  // namespace mk
