// Electron thermodynamics, synchrotron coefficients and the transfer scans as stand-alone kernels
// (API-parity path of /root/reference/mahakala/electrons.py:32-50 and transfer.py:30-144).
#include "common.cuh"
#include "ks_metric.cuh"
#include "snapshot.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

static_assert(sizeof(mk_emission_params) == sizeof(EmissionParams), "ABI struct must mirror EmissionParams");

__global__ void rlow_rhigh_kernel(const double* __restrict__ dens, const double* __restrict__ u,
                                  const double* __restrict__ beta, long n, double r_low, double r_high,
                                  double eg, double ig, double CL, double MP, double ME, double* __restrict__ out)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double b2 = beta[i] * beta[i];
        double T_ratio = (r_high * b2 + r_low) / (1. + b2);
        double t_e = (CL * CL) * (MP * u[i] * (eg - 1.) * (ig - 1.));
        t_e /= dens[i] * ((ig - 1.) + (eg - 1.) * T_ratio);
        out[i] = t_e / (ME * CL * CL);
    }
}

__global__ void synchrotron_kernel(EmissionParams P, const double* __restrict__ Ne, const double* __restrict__ Th,
                                   const double* __restrict__ B, const double* __restrict__ pitch,
                                   const double* __restrict__ nu, long n, int invariant, double rescale,
                                   double* __restrict__ em, double* __restrict__ ab)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double e, a;
        synchrotron(P, Ne[i], Th[i], B[i], sin(pitch[i]), nu[i], invariant, rescale, e, a);
        em[i] = e;
        ab[i] = a;
    }
}

// transfer.py:106-119: I += -dt[i-1] L (em[i] - ab[i] I) for i = nrows-1 .. 1; one thread per pixel
__global__ void solve_intensity_kernel(const double* __restrict__ em, const double* __restrict__ ab,
                                       const double* __restrict__ dt, long nrows, long npx, double L,
                                       double* __restrict__ I_out, double* __restrict__ dIs)
{
    long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= npx) return;
    double I = 0.0;
    long k = 0;
    for (long i = nrows - 1; i >= 1; i--, k++) {
        double dI = __dmul_rn(__dmul_rn(-dt[(i - 1) * npx + p], L), __dsub_rn(em[i * npx + p], __dmul_rn(ab[i * npx + p], I)));
        I = __dadd_rn(I, dI);
        if (dIs) dIs[k * npx + p] = dI;
    }
    I_out[p] = I;
}

// transfer.py:137-144
__global__ void attenuated_kernel(const double* __restrict__ em, const double* __restrict__ ab,
                                  const double* __restrict__ dt, long nrows, long npx, double L,
                                  double* __restrict__ out)
{
    long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= npx) return;
    double tau = 0.0;
    for (long i = 1; i < nrows; i++) {
        double d = dt[(i - 1) * npx + p];
        double src = __dmul_rn(__dmul_rn(-em[i * npx + p], d), L);
        double dtau = __dmul_rn(__dmul_rn(ab[i * npx + p], d), L);
        out[(i - 1) * npx + p] = exp(-tau) * src;
        tau = tau - dtau;
    }
}

// images.py:84-118 for one (state, primitives) pair in IEEE arithmetic with the reference's operation order: fluid
// frame -> beta, sigma, Theta_e, units -> j, alpha (invariant) -> sigma cut.  prims in canonical order.
__device__ __forceinline__ void emission_ieee(const EmissionParams& P, const KerrSchild& g, const double s[8],
                                              const double prims[8], double nu_obs, double& e, double& a)
{
    const double cos_fallback = 0.5000000000000001;      // cos(pi/3) as NumPy evaluates it
    double f, l[4];
    l[0] = 1.0;
    g.fl(s, f, l[1], l[2], l[3]);
    FluidScalars fs = fluid_frame(f, l, s, prims, cos_fallback);
    double Ne, Th, Bg, sigma;
    plasma_state(P, fs, Ne, Th, Bg, sigma);
    // the reference takes sin(arccos(c)) of the clamped cosine (transfer.py:62 after athenak.py:789-792)
    synchrotron(P, Ne, Th, Bg, sin(acos(fs.cos_pitch)), -fs.kdotu * nu_obs, 1, 1.0 / nu_obs, e, a);
    if (sigma > P.sigma_cut) { e = 0.0; a = 0.0; }
}

__global__ void emission_from_states_kernel(SnapshotView sn, EmissionParams P, KerrSchild g,
                                            const double* __restrict__ S, long n, double nu_obs,
                                            double* __restrict__ em, double* __restrict__ ab)
{
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        double s[8], prims[8];
#pragma unroll
        for (int m = 0; m < 8; m++) s[m] = S[p * 8 + m];
        double e = 0.0, a = 0.0;
        if (interp_prims(sn, s, prims)) emission_ieee(P, g, s, prims, nu_obs, e, a);
        em[p] = e;
        ab[p] = a;
    }
}

// Probe of the emission code of the fused render kernel on arbitrary (state, primitives) pairs: FAST = the
// emission_fast<NF> path exactly as render_kernel<NF> calls it (f, l from the point cache of the state), otherwise
// the IEEE chain above, frequency by frequency.
struct ProbeFreq { double nu[8], inv_nu[8]; };

template <int NF, bool FAST>
__global__ void emission_probe_kernel(EmissionParams P, EmissionConsts C, KerrSchild g, ProbeFreq F,
                                      const double* __restrict__ S, const double* __restrict__ prims_in, long n,
                                      double* __restrict__ em, double* __restrict__ ab)
{
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        double s[8], prims[8];
#pragma unroll
        for (int m = 0; m < 8; m++) { s[m] = S[p * 8 + m]; prims[m] = prims_in[p * 8 + m]; }
        if (FAST) {
            KerrSchild::Cache cache;
            g.radius(s, cache);
            double f, l[4];
            l[0] = 1.0;
            g.fl(s, cache, f, l[1], l[2], l[3]);
            emission_fast<NF>(P, C, f, l, s, prims, F.nu, F.inv_nu,
                              [&](int fq, double e, double a) { em[(long)fq * n + p] = e; ab[(long)fq * n + p] = a; });
        } else {
#pragma unroll
            for (int fq = 0; fq < NF; fq++) {
                double e, a;
                emission_ieee(P, g, s, prims, F.nu[fq], e, a);
                em[(long)fq * n + p] = e;
                ab[(long)fq * n + p] = a;
            }
        }
    }
}

static unsigned grid1d(long n, int threads)
{
    long blocks = (n + threads - 1) / threads;
    long cap = (long)sm_count() * 16;
    return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace mk
using namespace mk;

extern "C" int mk_rlow_rhigh(const double* dens, const double* u, const double* beta, long n, double r_low,
                             double r_high, double eg, double ig, double CL, double MP, double ME,
                             double* theta_e, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(dens && u && beta && theta_e, "null pointer");
    rlow_rhigh_kernel<<<grid1d(n, 256), 256, 0, (cudaStream_t)stream>>>(dens, u, beta, n, r_low, r_high, eg, ig, CL, MP, ME, theta_e);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_synchrotron(const mk_emission_params* c, const double* Ne, const double* theta_e,
                              const double* B, const double* pitch, const double* nu, long n, int invariant,
                              double rescale_nu, double* em, double* ab, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(c && Ne && theta_e && B && pitch && nu && em && ab, "null pointer");
    EmissionParams P;
    memcpy(&P, c, sizeof P);
    synchrotron_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(P, Ne, theta_e, B, pitch, nu, n, invariant, rescale_nu, em, ab);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_solve_specific_intensity(const double* em, const double* ab, const double* dt, long nrows,
                                           long npx, double L_unit, double* I_nu, double* dIs, void* stream)
{
    if (npx <= 0) return 0;
    MK_REQUIRE(em && ab && dt && I_nu, "null pointer");
    solve_intensity_kernel<<<(unsigned)((npx + 127) / 128), 128, 0, (cudaStream_t)stream>>>(em, ab, dt, nrows, npx, L_unit, I_nu, dIs);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_solve_attenuated_emissivity(const double* em, const double* ab, const double* dt, long nrows,
                                              long npx, double L_unit, double* out, void* stream)
{
    if (npx <= 0 || nrows <= 1) return 0;
    MK_REQUIRE(em && ab && dt && out, "null pointer");
    attenuated_kernel<<<(unsigned)((npx + 127) / 128), 128, 0, (cudaStream_t)stream>>>(em, ab, dt, nrows, npx, L_unit, out);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_emission_from_states(const mk_snapshot* snap, const mk_emission_params* params, double bhspin,
                                       const double* S, long n, double nu_obs, double* em, double* ab, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(snap && params && S && em && ab, "null pointer");
    EmissionParams P;
    memcpy(&P, params, sizeof P);
    KerrSchild g; g.set_spin(bhspin);
    emission_from_states_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(snap->view, P, g, S, n, nu_obs, em, ab);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int NF>
static void launch_probe(bool fast, const EmissionParams& P, const EmissionConsts& C, const KerrSchild& g,
                         const ProbeFreq& F, const double* S, const double* prims, long n, double* em, double* ab,
                         cudaStream_t stream)
{
    if (fast) emission_probe_kernel<NF, true><<<grid1d(n, 128), 128, 0, stream>>>(P, C, g, F, S, prims, n, em, ab);
    else emission_probe_kernel<NF, false><<<grid1d(n, 128), 128, 0, stream>>>(P, C, g, F, S, prims, n, em, ab);
}

extern "C" int mk_emission_probe(const mk_emission_params* params, double bhspin, const double* S,
                                 const double* prims, long n, int nfreq, const double* nu_obs, int fast, double* em,
                                 double* ab, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n <= 0) return 0;
    MK_REQUIRE(params && S && prims && nu_obs && em && ab, "null pointer");
    MK_REQUIRE(nfreq >= 1 && nfreq <= 8, "nfreq must be in 1..8");
    EmissionParams P;
    memcpy(&P, params, sizeof P);
    EmissionConsts C = make_emission_consts(P, nu_obs, nfreq);
    KerrSchild g; g.set_spin(bhspin);
    ProbeFreq F;
    for (int f = 0; f < 8; f++) { F.nu[f] = nu_obs[f < nfreq ? f : nfreq - 1]; F.inv_nu[f] = 1.0 / F.nu[f]; }
    switch (nfreq) {
        case 1: launch_probe<1>(fast, P, C, g, F, S, prims, n, em, ab, stream); break;
        case 2: launch_probe<2>(fast, P, C, g, F, S, prims, n, em, ab, stream); break;
        case 3: launch_probe<3>(fast, P, C, g, F, S, prims, n, em, ab, stream); break;
        case 4: launch_probe<4>(fast, P, C, g, F, S, prims, n, em, ab, stream); break;
        case 5: launch_probe<5>(fast, P, C, g, F, S, prims, n, em, ab, stream); break;
        case 6: launch_probe<6>(fast, P, C, g, F, S, prims, n, em, ab, stream); break;
        case 7: launch_probe<7>(fast, P, C, g, F, S, prims, n, em, ab, stream); break;
        default: launch_probe<8>(fast, P, C, g, F, S, prims, n, em, ab, stream); break;
    }
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}
