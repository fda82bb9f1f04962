/*
 * mahakala_b200 — C ABI of the B200-native per-ray hot path (libmahakala_b200.so).
 *
 * The reference (liamedeiros/Mahakala) is pure Python on JAX and has no FFI of its own; the interface a
 * replacement has to honour is the Python call surface of mahakala.geodesics / .transfer / .images and
 * the GRMHD fluid-model duck type.  Each entry point below names the reference function it replaces
 * (paths relative to /root/reference/mahakala/).  The Python mirror of that call surface lives in the
 * package `mahakala_b200` and binds these symbols with ctypes (mahakala_b200/_cabi.py); INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; mk_last_error_string() explains it;
 *   - all array arguments are DEVICE pointers (float64 unless noted), caller-owned, C-contiguous, with
 *     the shapes given; the library never allocates result buffers behind the caller's back;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous with
 *     respect to the host unless stated otherwise;
 *   - one CUDA context per device; call from the host thread that made the device current.
 */
#ifndef MAHAKALA_B200_H
#define MAHAKALA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MK_ABI_VERSION 1

/* built-in spacetimes (metric plugins compiled into the library) */
#define MK_METRIC_KERR_SCHILD 0      /* closed-form Cartesian Kerr-Schild: geodesics.py:88-104 */
#define MK_METRIC_KERR_SCHILD_DUAL 1 /* same metric through the generic dual-number plugin path */
#define MK_METRIC_KERR_SCHILD_STRICT 2 /* literal jets + 4x4 inverse in non-contracted IEEE arithmetic: bit-identical to
                                        the CPU restatement of geodesics.py:233-351; mk_integrate final-state mode only */
#define MK_METRIC_PLUGIN_BASE 16     /* ids >= 16: spacetimes registered at run time (mk_register_metric) */

/* ---- library ------------------------------------------------------------------------------- */
int mk_abi_version(void);
const char* mk_last_error_string(void);
/* multiprocessor count, memory clock etc. of the current device; any pointer may be NULL */
int mk_device_info(int* sm_count, int* cc_major, int* cc_minor, int* sm_clock_khz, long* total_mem_bytes);
/* DFMA microbenchmark: measured FP64 FMA throughput (TFLOP/s, 2 flop per FMA) of the current device.
   Synchronous.  Used as the roofline denominator of the integrator. */
int mk_measure_fp64_peak(int iters, double* tflops_out, double* ms_out);
/* evaluates the kernels' MUFU-seeded reciprocal, square root and reciprocal square root on x (n,) so that
   their accuracy can be checked against IEEE results (positive normal inputs) */
int mk_fast_math_probe(const double* x, long n, double* rcp, double* sqrt_out, double* rsqrt_out, void* stream);
/* the fused render kernel's exp(-x) (x >= 0) and cube root / reciprocal cube root (positive x inside the float
   range) evaluated on x (n,), for accuracy checks against libm */
int mk_transcendental_probe(const double* x, long n, double* exp_neg, double* cbrt_out, double* rcbrt_out,
                            void* stream);

/* ---- user-registered spacetimes ------------------------------------------------------------------ */
/*
 * The reference lets a user change spacetime by replacing the module-level metric()/imetric()
 * (geodesics.py:88-104, :304-305, :339-347); jax.jacfwd supplies the derivatives.  Here `source` is CUDA C++
 * defining `struct UserMetric` (contract in mahakala_b200/csrc/plugin_tu.cuh): the covariant metric on a
 * generic scalar type, the step-rule radius and the horizon radius.  It is compiled with NVRTC for sm_100a
 * against the headers in include_dir (the package's csrc directory); derivatives come from forward-mode dual
 * numbers and the inverse from a 4x4 adjugate.  The returned id (>= MK_METRIC_PLUGIN_BASE) is accepted by
 * mk_integrate, mk_integrate_paged, mk_rhs, mk_rk4_step, mk_metric and mk_initial_condition_metric; the
 * `bhspin` argument of those calls becomes UserMetric::params[0], params[1..7] come from
 * mk_metric_set_params.  Compilation needs no GPU.  log (optional) receives the compiler log.
 */
int mk_register_metric(const char* name, const char* source, const char* include_dir, int* metric_id,
                       char* log, long log_capacity);
int mk_metric_set_params(int metric_id, const double* params8);
/* geodesics.py:219-230 initial_condition with the selected spacetime (built-in ids use Kerr-Schild) */
int mk_initial_condition_metric(int metric_id, double bhspin, const double* s0_x, const double* s0_v, long n,
                                double* s0, void* stream);

/* ---- camera: geodesics.py:29-55 initialize_geodesics_at_camera ------------------------------- */
/* 'grid' camera (geodesics.py:158-181 + :219-230): s0 (n*n, 8) = [t,x,y,z,k^t,k^x,k^y,k^z], pixel
   index ix*n+iy.  cos_i/sin_i are cos/sin of the inclination in radians, evaluated by the caller.
   nullify = 0 returns the raw positions and un-normalised directions of get_initial_grid (:137-181). */
int mk_camera_grid(double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
                   double fov_upper, long pixels_per_side, int nullify, double* s0, void* stream);
/* arbitrary image-plane points (geodesics.py:107-134 get_camera_pixel + :219-230): x_img, y_img (n,) */
int mk_camera_points(double bhspin, double cos_i, double sin_i, double distance, const double* x_img,
                     const double* y_img, long n, int nullify, double* s0, void* stream);
/* geodesics.py:219-230 initial_condition: s0_x, s0_v (4, n) -> s0 (n, 8), spatial k rescaled to null */
int mk_initial_condition(double bhspin, const double* s0_x, const double* s0_v, long n, double* s0,
                         void* stream);

/* ---- geodesics: geodesics.py:233-281 geodesic_integrator, :370-378 last-point rule ------------- */
/*
 * Integrates npx rays for at most N iterations with the fixed rule dt = -(r - r_H)/div.
 *   final_state (npx, 8)  state at which each ray froze (or after N accepted steps)      [optional]
 *   nsteps      (npx,) i32 accepted steps n (rows with dt != 0)                           [optional]
 *   r_last      (npx,)    radius_cal(S[argmax(dt) - 1]) incl. the reference's index wrap  [optional]
 *   S (nrows, npx, 8), dt (nrows, npx): trajectory dump; rows 0..min(n, nrows-1) of each ray are
 *       written (row i = state before step i, row n = frozen state with dt 0); call
 *       mk_fill_frozen_rows afterwards to replicate the frozen row into rows n+1..nrows-1.   [optional]
 *   total_steps: device counter incremented by the sum of n over all rays                 [optional]
 */
int mk_integrate(int metric_id, double bhspin, long N, long npx, const double* s0, double div, double tol,
                 double* final_state, int32_t* nsteps, double* r_last, double* S, double* dt, long nrows,
                 unsigned long long* total_steps, void* stream);
/* OPTIONAL integrator, not in the reference: embedded Dormand-Prince 5(4) with step-size
   control (csrc/adaptive.cuh) instead of classical RK4 under the fixed rule of geodesics.py:246-269.  Same termination
   test (a ray lives while tol <= radius - r_H <= 1500), same freeze semantics (a step that would end outside that
   range is rejected), h < 0, and |h| <= cap (radius - r_H).  Per-step error bound atol + rtol |y| on position and
   wavevector.  metric_id: MK_METRIC_KERR_SCHILD, MK_METRIC_KERR_SCHILD_DUAL or a registered metric (its radius() and
   horizon() only decide when a ray has ended; the step size no longer depends on them).
     final_state (npx, 8), nsteps (npx,) accepted steps, nrejected (npx,) rejected trial steps, r_last (npx,) radius of
     the final state (~r_H + tol: captured; hundreds of M: escaped) -- all optional. */
int mk_integrate_adaptive(int metric_id, double bhspin, long N, long npx, const double* s0, double rtol, double atol,
                          double tol, double cap, double* final_state, int32_t* nsteps, int32_t* nrejected,
                          double* r_last, void* stream);
int mk_fill_frozen_rows(double* S, double* dt, const double* final_state, const int32_t* nsteps, long npx,
                        long nrows, void* stream);
/*
 * Single-pass trajectory dump into a paged (ragged) store: no row count has to be known in advance and no
 * padding is written.  Every warp of the persistent kernel appends to its own log: per loop iteration one
 * SLOT holding the rows of its 32 lanes (32 x 64 B of state, contiguous, then 32 step sizes), so all dump
 * stores are fully coalesced.  Page p occupies 4608 doubles at pages + 4608*p: [16 slots][32 lanes][8] states
 * followed by [16][32] step sizes (mk_page_rows() == 16 slots per page); page_next (max_pages,) chains the
 * pages of one warp (-1 terminates).  Lanes are refilled when their ray freezes, so a ray is the column
 * `lane` of consecutive slots starting at page_first[ray] = {page, slot*32 + lane} (page_first is (npx, 2)).
 * A ray stores rows 0..n (row n = frozen state, dt = 0; N rows when it never froze).  page_counter (one u32,
 * zeroed by the caller) counts pages handed out; *overflow becomes 1 if more than max_pages were needed
 * (results other than the trajectories stay valid).  mk_paged_gather materialises the reference layout
 * S (nrows, nsel, 8), dt (nrows, nsel) for the rays ray_idx[0..nsel) (NULL = rays 0..nsel-1).
 */
int mk_page_rows(void);
int mk_integrate_paged(int metric_id, double bhspin, long N, long npx, const double* s0, double div,
                       double tol, double* final_state, int32_t* nsteps, double* r_last, double* pages,
                       int32_t* page_next, int32_t* page_first, unsigned int* page_counter, long max_pages,
                       int32_t* overflow, unsigned long long* total_steps, void* stream);
/*
 * The same launch as one PARTICIPANT of a multi-GPU job (BASELINE north_star: "tiles handed out across the GPUs from
 * a dynamic queue ... gathered at the end").  Every GPU of the node runs this call on the same (npx, 8) bundle:
 *   queue        one zero-initialised u32 in the memory of ONE GPU, mapped by all others (mk_ipc_alloc / mk_ipc_open);
 *                warps take chunks of up to 32 queue positions with system-scope atomics over NVLink and request
 *                the next chunk while the current one is being integrated, so the round trip is never waited for;
 *                chunks shrink to single rays as the queue empties (participants = number of GPUs on the queue);
 *   ray_order    (npx,) int32 or NULL: queue position -> ray index.  Handing out the long rays (photon ring) first
 *                keeps the tail of the queue short; results stay indexed by ray;
 *   final_state, nsteps, r_last, page_first
 *                may point into the gathering GPU's memory: each ray's results are stored there by whichever GPU
 *                integrated it (the gather is fused into the kernel, no collective on the data path);
 *   pages, page_next, page_counter, overflow, total_steps
 *                stay LOCAL: trajectories are logged where they are computed; page_first records
 *                page_id_offset + local page (callers use rank * max_pages) so the owner of a ray can be found.
 * Results are identical to a single-GPU launch ray for ray (per-ray arithmetic does not depend on scheduling).
 */
int mk_integrate_shared(int metric_id, double bhspin, long N, long npx, const double* s0, double div, double tol,
                        double* final_state, int32_t* nsteps, double* r_last, double* pages, int32_t* page_next,
                        int32_t* page_first, unsigned int* page_counter, long max_pages, int32_t* overflow,
                        unsigned long long* total_steps, unsigned int* queue, const int32_t* ray_order,
                        long page_id_offset, int participants, void* stream);
int mk_paged_gather(const double* pages, const int32_t* page_next, const int32_t* page_first,
                    const int32_t* nsteps, const long* ray_idx, long nsel, long nrows, long N, double* S,
                    double* dt, void* stream);
/* geodesics.py:405-435 find_shadow_bisection_angles in one launch (built-in Kerr-Schild spacetime): for each of the n
   image-plane angles (cos_angle, sin_angle: DEVICE arrays of cos / sin evaluated by the caller) n_iter bisection
   steps on the bracket [inner0, outer0]: ray through the mid radius (geodesics.py:107-134 + :219-230), integrated
   for at most N steps with div / tol (:354-378), bracket halved on "classifier radius < limit".  inner_out /
   outer_out (n,) receive the final brackets; the reference returns inner. */
int mk_shadow_bisection(double bhspin, double cos_i, double sin_i, double distance, const double* cos_angle,
                        const double* sin_angle, long n, long N, double div, double tol, int n_iter, double inner0,
                        double outer0, double limit, double* inner_out, double* outer_out, void* stream);
/* geodesics.py:284-291 radius_cal for n points of stride `stride` doubles (x at offsets 1..3) */
int mk_radius_cal(double bhspin, const double* x, long n, long stride, double* r, void* stream);
/* geodesics.py:294-314 rhs on a bundle: state (n, 8) -> (n, 8); metric_id selects the plugin */
int mk_rhs(int metric_id, double bhspin, const double* state, long n, double* out, void* stream);
/* geodesics.py:317-336 RK4_gen: one RK4 step of a bundle with per-ray dt (n,) */
int mk_rk4_step(int metric_id, double bhspin, const double* state, const double* dt, long n, double* out,
                void* stream);
/* geodesics.py:88-104 metric and :339-347 imetric at n points x (n, 4): g, gi (n, 4, 4); either may be NULL */
int mk_metric(int metric_id, double bhspin, const double* x, long n, double* g, double* gi, void* stream);

/* ---- GRMHD snapshot: grmhd/athenak.py AthenakFluidModel ------------------------------------- */
typedef struct mk_snapshot mk_snapshot;   /* opaque; owns the repacked, device-resident snapshot */

/*
 * Repack a ghost-padded AthenaK-style snapshot for sampling (replaces the per-call upload of
 * athenak.py:693).  Inputs are DEVICE arrays:
 *   meshblocks (nmb, 8, nk+2, nj+2, ni+2)   the reference's self.all_meshblocks (athenak.py:105-158)
 *   geom (nmb, 12)  per block: x1f[0], x2f[0], x3f[0], x1f[-1], x2f[-1], x3f[-1], x1v[0], x2v[0], x3v[0],
 *                         dx1, dx2, dx3   (dx = x_v[1] - x_v[0], athenak.py:686-691)
 *   grid (gn[2], gn[1], gn[0]) int32 block-lookup table over the bounding box, or NULL for the
 *        reference's linear scan (athenak.py:663-670); cell (c0,c1,c2) covers
 *        g0[d] + c_d / ginv[d] .. in dimension d.
 * Host arrays: prim_index[8] = index in `meshblocks` of dens, eint, velx, vely, velz, bcc1, bcc2, bcc3
 * (athenak.py:697-710); gn, g0, ginv, bbox_lo, bbox_hi (3 each).
 * store_f32 != 0 stores cells as float32 (caller guarantees the values are float32-representable).
 * meshblocks == NULL allocates the snapshot without filling the cells (a replica that receives them from
 * another rank through mk_snapshot_cells + an NCCL broadcast).
 * The call is synchronous with respect to `stream` on return of the handle (the inputs may be freed).
 */
int mk_snapshot_create(long nmb, long nk, long nj, long ni, const double* meshblocks, const int* prim_index,
                       const double* geom, const int* grid, const int* gn, const double* g0,
                       const double* ginv, const double* bbox_lo, const double* bbox_hi, int store_f32,
                       mk_snapshot** out, void* stream);
/*
 * Same snapshot, built straight from the arrays an AthenaK .athdf file holds: the ghost-zone fill of the
 * reference's loader (athenak.py:105-158 layout, :208-229 same-level copy, :231-514 refinement boundaries:
 * coarser neighbour -> injection, finer neighbour -> mean of the 8 covered cells, domain boundary -> 0) runs
 * on the GPU fused with the repack, so no ghost-padded host array and no (nmb, 8, nk+2, nj+2, ni+2) upload
 * is needed.  DEVICE arrays: uov (n_uov, nmb, nk, nj, ni) and B (n_B, nmb, nk, nj, ni), n_uov + n_B == 8,
 * float32 (src_f32 != 0, what AthenaK writes) or float64; geom / grid as above.  HOST arrays:
 * logical_locations (nmb, 3) int32 = (li, lj, lk) and levels (nmb) int32 (athenak.py:160-206), prim_index[8] =
 * position of dens, eint, velx, vely, velz, bcc1, bcc2, bcc3 in the concatenation [uov..., B...].
 * store_mode: 0 = float64 cells, 1 = float32 cells, 2 = float32 when every stored value (ghost averages
 * included) is float32-representable, else float64; *stored_f32 (optional) reports the choice.
 * Synchronous with respect to `stream` on return.
 */
int mk_snapshot_create_from_interiors(long nmb, long nk, long nj, long ni, const void* uov, int n_uov,
                                      const void* B, int n_B, int src_f32, const int* prim_index,
                                      const int* logical_locations, const int* levels, const double* geom,
                                      const int* grid, const int* gn, const double* g0, const double* ginv,
                                      const double* bbox_lo, const double* bbox_hi, int store_mode,
                                      int* stored_f32, mk_snapshot** out, void* stream);
/* Host -> device copy of a large PAGEABLE host array (the interior arrays a loader returns) through two pinned
   staging buffers: host_threads (0 = up to 8) threads fill one buffer while the DMA engine empties the other;
   host_threads < 0 pins the caller's pages in place instead (cudaHostRegister) and lets the DMA engine read them
   directly -- no host-side copy -- falling back to staging when they cannot be pinned.
   Synchronous with respect to the host on return (the source may be freed); ordered on `stream`. */
int mk_upload_pageable(void* dst_device, const void* src_host, long bytes, int host_threads, void* stream);
/* inverse of the repack: cells -> the reference's self.all_meshblocks (nmb, 8, nk+2, nj+2, ni+2) float64
   (device), primitive q of the canonical order written to position prim_index[q] */
int mk_snapshot_unpack(const mk_snapshot* snap, const int* prim_index, double* meshblocks, void* stream);
/* Analytic fluid source instead of snapshot cells (BASELINE cfg3: Keplerian thin torus, power-law density,
   toroidal field at fixed beta; not in the reference).  params9 (host) = {fluid_gamma, R0,
   R_in, p, h, u0, beta0, dens_scale, r_out}.  The handle works with every mk_sample_* / mk_render call. */
int mk_snapshot_create_torus(const double* params9, mk_snapshot** out);
int mk_snapshot_destroy(mk_snapshot* snap);
/* bytes of HBM held by the snapshot */
long mk_snapshot_bytes(const mk_snapshot* snap);
/* raw device pointer + byte size of the repacked cell array (for NCCL broadcast of a replicated
   snapshot across ranks: every rank creates the same-shaped snapshot, rank 0 broadcasts the cells) */
int mk_snapshot_cells(mk_snapshot* snap, void** cells, long* bytes);

/* athenak.py:639-812 get_fluid_scalars_from_geodesics: S (n, 8) -> out (5, n) rows dens, u,
   pitch_angle, kdotu, b.  Kerr-Schild spacetime of spin bhspin. */
int mk_sample_scalars(const mk_snapshot* snap, double bhspin, const double* S, long n,
                      double fallback_pitch_angle, double* out, void* stream);
/* athenak.py:527-637 get_prims_from_geodesics: S (n, 8) -> out (8, n) rows dens, u, U1..3, B1..3 */
int mk_sample_prims(const mk_snapshot* snap, const double* S, long n, double* out, void* stream);

/* ---- thermodynamics and transfer: electrons.py, transfer.py, images.py ----------------------- */
typedef struct mk_emission_params {
    double fluid_gamma, r_low, r_high, electron_gamma, ion_gamma;   /* electrons.py:32-33 */
    double Ne_unit, B_unit, L_unit;                                   /* grmhd/grmhd.py:40-55 */
    double sigma_cut;                                                 /* images.py:116 */
    double EE, CL, ME, MP, HPL;                                       /* constants.py:23-31 */
    double two_11_12;                                                 /* 2**(11/12), transfer.py:65 */
} mk_emission_params;

/* electrons.py:32-50 rlow_rhigh_model, elementwise over n values */
int mk_rlow_rhigh(const double* dens, const double* u, const double* beta, long n, double r_low,
                  double r_high, double electron_gamma, double ion_gamma, double CL, double MP, double ME,
                  double* theta_e, void* stream);
/* transfer.py:30-86 synchrotron_coefficients, elementwise over n values; nu may be per-element */
int mk_synchrotron(const mk_emission_params* constants, const double* Ne, const double* theta_e,
                   const double* B, const double* pitch_angle, const double* nu, long n, int invariant,
                   double rescale_nu, double* emissivity, double* absorptivity, void* stream);
/* transfer.py:89-119 solve_specific_intensity: em, ab (nrows, npx), dt (nrows, npx) -> I (npx,);
   dIs (nrows-1, npx) optional (dIs=True variant, rows in scan order i = nrows-1 .. 1) */
int mk_solve_specific_intensity(const double* em, const double* ab, const double* dt, long nrows, long npx,
                                double L_unit, double* I_nu, double* dIs, void* stream);
/* transfer.py:122-144 solve_attenuated_emissivity -> out (nrows-1, npx) */
int mk_solve_attenuated_emissivity(const double* em, const double* ab, const double* dt, long nrows,
                                   long npx, double L_unit, double* out, void* stream);
/* images.py:84-118 for one chunk: S (nrows*npx, 8) -> invariant em, ab (nrows*npx) at frequency nu_obs
   (fluid scalars -> beta, sigma, Theta_e, units -> j, alpha -> sigma cut), one fused elementwise kernel */
int mk_emission_from_states(const mk_snapshot* snap, const mk_emission_params* params, double bhspin,
                            const double* S, long n, double nu_obs, double* em, double* ab, void* stream);

/* Probe of the emission code on arbitrary inputs: S (n, 8) states, prims (n, 8) primitives in canonical order dens,
   eint, U1..3, B1..3 (no snapshot lookup) -> invariant em, ab (nfreq, n) at the HOST frequencies nu_obs[nfreq <= 8].
   fast != 0 runs emission_fast<nfreq>, the straight-line code of the fused render kernel; fast == 0 the literal IEEE
   chain of images.py:87-118 + athenak.py:760-794 + transfer.py:56-86.  Lets the tests hold the fused path to the
   reference's special cases (sigma cut, Theta_e floor, X limit, Planck series switch, pitch clamp, NaN -> 0). */
int mk_emission_probe(const mk_emission_params* params, double bhspin, const double* S, const double* prims, long n,
                      int nfreq, const double* nu_obs, int fast, double* em, double* ab, void* stream);

/* ---- fused render: images.py:30-144 make_image ------------------------------------------------ */
/*
 * One persistent kernel: camera ray -> RK4 geodesic -> snapshot sample -> j, alpha -> intensity, all in
 * registers; no trajectory is materialised.  Either a grid camera (s0 == NULL: res x res pixels,
 * pixel index ix*res+iy as images.py:144) or explicit rays s0 (npx, 8).
 *   image (nfreq, npx)  specific intensity per observing frequency nu_obs[f] (nfreq <= 8)
 *   nsteps (npx,) optional accepted steps; total_steps / total_samples optional device counters
 *   (sum of accepted steps; number of in-domain samples)
 *   queue: optional device counter (zero-initialised by the caller) from which warps pull 32-ray
 *   patches; it may live in a peer GPU's memory so that several GPUs share one dynamic tile queue.
 *   The launch processes patches patch_begin + k*patch_stride < patch_end (k from the queue): 0, -1, 1
 *   = everything; rank, -1, world = static interleaved sharding.  patch_order (device int32, one entry per
 *   patch, or NULL) permutes the order in which patches are handed out: putting the long rays near the
 *   photon ring first shortens the tail of the dynamic queue (longest-processing-time-first).  image may likewise be a peer pointer
 *   (tiles are written where the gather would put them).
 */
int mk_render(double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
              double fov_upper, long res, const double* s0, long npx, long N, double div, double tol,
              const mk_snapshot* snap, const mk_emission_params* params, int nfreq, const double* nu_obs,
              double* image, int32_t* nsteps, unsigned long long* total_steps,
              unsigned long long* total_samples, unsigned int* queue, long patch_begin, long patch_end,
              long patch_stride, const int* patch_order, void* stream);
/* The same contract through the warp-specialised LONG-PATCH kernel (csrc/render_pipeline.cuh): one CTA per patch, a
   producer warp integrates the geodesics and hands every state to three consumer warps (sample + emission) through a
   shared-memory ring, so that the snapshot sample leaves the critical path of the dependent RK4 steps (0.73 instead of
   1.95 us per step for a lone patch on B200).  Pixels are bit-identical to mk_render's.  Meant for the few hundred
   patches that contain photon-ring rays (patch_begin .. patch_end of a longest-first patch_order), launched on a
   high-priority stream next to an mk_render launch for the rest; nfreq <= 8 like mk_render.
   exclusive = 0: 128-thread CTAs (one patch each) that share their SM with other resident CTAs; exclusive = 1..4:
   512-thread CTAs that fill the register file of an SM, so that no warp of another launch competes with the producers
   for FP64 issue slots, with that many patch groups at work (one producer per SM sub-partition; the warps of the
   other groups exit at once).  max_ctas > 0 caps the grid (several GPUs
   pulling from one queue: about (patches / 4) / GPUs each, so that the long patches spread over all of them). */
int mk_render_long(double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
                   double fov_upper, long res, const double* s0, long npx, long N, double div, double tol,
                   const mk_snapshot* snap, const mk_emission_params* params, int nfreq, const double* nu_obs,
                   double* image, int32_t* nsteps, unsigned long long* total_steps,
                   unsigned long long* total_samples, unsigned int* queue, long patch_begin, long patch_end,
                   long patch_stride, const int* patch_order, int exclusive, int max_ctas, void* stream);
/* The same with a selectable spacetime: MK_METRIC_KERR_SCHILD (== mk_render) or a run-time registered metric
   (id >= MK_METRIC_PLUGIN_BASE): geodesics through the plugin's dual-number derivatives, fluid-frame algebra
   (athenak.py:760-786) with the plugin's own covariant / contravariant metric at every sample.  The reference obtains
   this by replacing its module-level metric (geodesics.py:88-104; athenak.py:34 imports it).  Registered spacetimes
   carry one frequency per launch (nfreq > 1 = one launch per frequency). */
int mk_render_metric(int metric_id, double bhspin, double cos_i, double sin_i, double distance, double fov_lower,
                     double fov_upper, long res, const double* s0, long npx, long N, double div, double tol,
                     const mk_snapshot* snap, const mk_emission_params* params, int nfreq, const double* nu_obs,
                     double* image, int32_t* nsteps, unsigned long long* total_steps,
                     unsigned long long* total_samples, unsigned int* queue, long patch_begin, long patch_end,
                     long patch_stride, const int* patch_order, void* stream);
/* number of 32-ray patches mk_render splits a job into (grid camera: 4x8 pixel patches) */
long mk_render_patch_count(long res, const double* s0, long npx);

/* ---- peer memory for the multi-GPU tile queue and in-kernel gather ----------------------------- */
/* cudaMalloc'ed, zero-filled buffer that other processes on the node can map: handle receives the 64-byte
   cudaIpcMemHandle_t to ship to the peers (e.g. with torch.distributed.broadcast_object_list). */
int mk_ipc_alloc(long bytes, void** ptr, unsigned char* handle64);
int mk_ipc_open(const unsigned char* handle64, void** ptr);
int mk_ipc_close(void* ptr);
int mk_ipc_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* MAHAKALA_B200_H */
