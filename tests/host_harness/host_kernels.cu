// TEST INFRASTRUCTURE: the arithmetic of the ray kernels compiled FOR THE HOST from the product's own headers
// (mahakala_b200/csrc/*.cuh are __host__ __device__), so that the CPU test suite can hold the closed-form
// acceleration, the RK4 step, the step rule and the fused emission chain to the oracle without a GPU.  Nothing in
// the product loads this library; the product path stays CUDA-only.  Built by tests/host_harness/build.py.
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>
#include "../../mahakala_b200/csrc/integrate_kernel.cuh"
#include "../../mahakala_b200/csrc/adaptive.cuh"
#include "../../mahakala_b200/csrc/ks_metric.cuh"
#include "../../mahakala_b200/csrc/sample.cuh"

using namespace mk;

static KerrSchild make_ks(double a)
{
    KerrSchild g;
    g.set_spin(a);
    return g;
}

extern "C" void hk_rhs(long n, const double* s, double a, double* out)
{
    KerrSchild g = make_ks(a);
    for (long i = 0; i < n; i++) {
        double acc[4];
        g.accel(s + 8 * i, s + 8 * i + 4, acc);
        for (int m = 0; m < 4; m++) { out[8 * i + m] = s[8 * i + 4 + m]; out[8 * i + 4 + m] = acc[m]; }
    }
}

extern "C" void hk_rk4(long n, const double* s, const double* dt, double a, double* out)
{
    KerrSchild g = make_ks(a);
    for (long i = 0; i < n; i++) rk4_step(g, s + 8 * i, dt[i], out + 8 * i);
}

// the per-ray loop of integrate_kernel.cuh (integrate_one: the lane logic of the kernel without the warp machinery)
extern "C" void hk_integrate(long n, const double* s0, int N, double div, double tol, double a, double* final_state,
                             int* nsteps, double* r_last)
{
    KerrSchild g = make_ks(a);
    StepRule rule;
    rule.div = div; rule.inv_div = 1.0 / div; rule.tol = tol; rule.rH = g.rH;
#pragma omp parallel for schedule(dynamic, 16)
    for (long p = 0; p < n; p++) {
        double s[8];
        std::memcpy(s, s0 + 8 * p, sizeof s);
        int it = 0;
        r_last[p] = integrate_one(g, rule, s, N, it);
        std::memcpy(final_state + 8 * p, s, sizeof s);
        nsteps[p] = it;
    }
}

// the per-ray loop of adaptive.cuh (embedded Dormand-Prince 5(4))
extern "C" void hk_integrate_adaptive(long n, const double* s0, int N, double rtol, double atol, double tol, double cap,
                                      double a, double* final_state, int* nsteps, int* nrejected, double* r_last)
{
    KerrSchild g = make_ks(a);
    AdaptiveRule R;
    R.rtol = rtol; R.atol = atol; R.tol = tol; R.far = 1500.0; R.rH = g.rH; R.cap = cap; R.div0 = 40.0;
#pragma omp parallel for schedule(dynamic, 16)
    for (long p = 0; p < n; p++) {
        double s[8];
        std::memcpy(s, s0 + 8 * p, sizeof s);
        int it = 0, rej = 0;
        r_last[p] = integrate_one_adaptive(g, R, s, N, it, rej);
        std::memcpy(final_state + 8 * p, s, sizeof s);
        nsteps[p] = it;
        nrejected[p] = rej;
    }
}

struct hk_params { double v[15]; };

extern "C" void hk_emission_fast(long n, const double* S, const double* prims, const double* params15, double a,
                                 int nfreq, const double* nu_obs, double* em, double* ab)
{
    KerrSchild g = make_ks(a);
    EmissionParams P;
    std::memcpy(&P, params15, sizeof P);
    EmissionConsts C = make_emission_consts(P, nu_obs, nfreq);
    double nu[8], inu[8];
    for (int f = 0; f < 8; f++) { nu[f] = nu_obs[f < nfreq ? f : nfreq - 1]; inu[f] = 1.0 / nu[f]; }
    for (long p = 0; p < n; p++) {
        const double* s = S + 8 * p;
        KerrSchild::Cache c;
        g.radius(s, c);
        double f, l[4];
        l[0] = 1.0;
        g.fl(s, c, f, l[1], l[2], l[3]);
        auto sink = [&](int fq, double e, double b) { em[(long)fq * n + p] = e; ab[(long)fq * n + p] = b; };
        switch (nfreq) {
            case 1: emission_fast<1>(P, C, f, l, s, prims + 8 * p, nu, inu, sink); break;
            case 2: emission_fast<2>(P, C, f, l, s, prims + 8 * p, nu, inu, sink); break;
            case 3: emission_fast<3>(P, C, f, l, s, prims + 8 * p, nu, inu, sink); break;
            case 4: emission_fast<4>(P, C, f, l, s, prims + 8 * p, nu, inu, sink); break;
            case 5: emission_fast<5>(P, C, f, l, s, prims + 8 * p, nu, inu, sink); break;
            case 6: emission_fast<6>(P, C, f, l, s, prims + 8 * p, nu, inu, sink); break;
            case 7: emission_fast<7>(P, C, f, l, s, prims + 8 * p, nu, inu, sink); break;
            default: emission_fast<8>(P, C, f, l, s, prims + 8 * p, nu, inu, sink); break;
        }
    }
}

extern "C" void hk_rhs_v1(long n, const double* s, double a, double* out)
{
    KerrSchild g = make_ks(a);
    for (long i = 0; i < n; i++) {
        double acc[4];
        g.accel_v1(s + 8 * i, s + 8 * i + 4, acc);
        for (int m = 0; m < 4; m++) { out[8 * i + m] = s[8 * i + 4 + m]; out[8 * i + 4 + m] = acc[m]; }
    }
}

// Sampling path (block lookup, cell index, trilinear gather) on HOST arrays laid out like the device snapshot:
// cells [mb][k][j][i][8] (canonical primitive order, ghost padded), geom (nmb, 16), int32 block grid.
extern "C" void hk_sample_prims(int kind, const void* cells, int is_f32, int nmb, int nk, int nj, int ni,
                                const double* geom16, const int* grid, const int* gn, const double* g0,
                                const double* ginv, const double* bbox_lo, const double* bbox_hi, int dx_pow2,
                                long n, const double* S, double* out)
{
    SnapshotView v;
    std::memset(&v, 0, sizeof v);
    v.source = 0; v.cells = cells; v.is_f32 = is_f32;
    v.nmb = nmb; v.nk = nk; v.nj = nj; v.ni = ni;
    v.sj = (long)(ni + 2) * 8; v.sk = v.sj * (nj + 2); v.sb = v.sk * (nk + 2);
    v.geom = geom16; v.grid = grid; v.dx_pow2 = dx_pow2;
    for (int d = 0; d < 3; d++) {
        v.gn[d] = grid ? gn[d] : 0; v.g0[d] = grid ? g0[d] : 0; v.ginv[d] = grid ? ginv[d] : 0;
        v.bbox_lo[d] = bbox_lo[d]; v.bbox_hi[d] = bbox_hi[d];
    }
    for (long p = 0; p < n; p++) {
        double prims[8];
        if (kind == SNAP_F64_GRID_POW2) interp_prims_kind<SNAP_F64_GRID_POW2>(v, S + 8 * p, prims);
        else if (kind == SNAP_F32_GRID_POW2) interp_prims_kind<SNAP_F32_GRID_POW2>(v, S + 8 * p, prims);
        else interp_prims(v, S + 8 * p, prims);
        for (int q = 0; q < 8; q++) out[(long)q * n + p] = prims[q];
    }
}
