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

/* ---- library ------------------------------------------------------------------------------- */
int mk_abi_version(void);
const char* mk_last_error_string(void);
/* multiprocessor count, memory clock etc. of the current device; any pointer may be NULL */
int mk_device_info(int* sm_count, int* cc_major, int* cc_minor, int* sm_clock_khz, long* total_mem_bytes);
/* DFMA microbenchmark: measured FP64 FMA throughput (TFLOP/s, 2 flop per FMA) of the current device.
   Synchronous.  Used as the roofline denominator of the integrator. */
int mk_measure_fp64_peak(int iters, double* tflops_out, double* ms_out);

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
int mk_fill_frozen_rows(double* S, double* dt, const double* final_state, const int32_t* nsteps, long npx,
                        long nrows, void* stream);
/* geodesics.py:284-291 radius_cal for n points of stride `stride` doubles (x at offsets 1..3) */
int mk_radius_cal(double bhspin, const double* x, long n, long stride, double* r, void* stream);
/* geodesics.py:294-314 rhs on a bundle: state (n, 8) -> (n, 8); metric_id selects the plugin */
int mk_rhs(int metric_id, double bhspin, const double* state, long n, double* out, void* stream);
/* geodesics.py:317-336 RK4_gen: one RK4 step of a bundle with per-ray dt (n,) */
int mk_rk4_step(int metric_id, double bhspin, const double* state, const double* dt, long n, double* out,
                void* stream);
/* geodesics.py:88-104 metric and :339-347 imetric at n points x (n, 4): g, gi (n, 4, 4); either may be NULL */
int mk_metric(int metric_id, double bhspin, const double* x, long n, double* g, double* gi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAHAKALA_B200_H */
