// Snapshot repack + stand-alone sampling kernels (the API-parity path of
// /root/reference/mahakala/grmhd/athenak.py:527-812; the fused path lives in render.cu).
#include "common.cuh"
#include "ks_metric.cuh"
#include "snapshot.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

struct PrimIndex { int p[8]; };

// reference layout (nmb, 8, nk+2, nj+2, ni+2) -> cells[mb][k][j][i][8] in canonical primitive order
template <class CellT>
__global__ void repack_kernel(const double* __restrict__ src, CellT* __restrict__ dst, long nmb, long cells_per_block,
                              PrimIndex pi)
{
    long total = nmb * cells_per_block;
    for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < total; c += (long)gridDim.x * blockDim.x) {
        long mb = c / cells_per_block, w = c - mb * cells_per_block;
        const double* s = src + mb * 8 * cells_per_block + w;
        CellT v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = (CellT)s[(long)pi.p[q] * cells_per_block];
#pragma unroll
        for (int q = 0; q < 8; q++) dst[c * 8 + q] = v[q];
    }
}

__global__ void sample_scalars_kernel(SnapshotView sn, KerrSchild g, const double* __restrict__ S, long n,
                                      double cos_fallback, double* __restrict__ out)
{
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        double s[8], prims[8];
#pragma unroll
        for (int m = 0; m < 8; m++) s[m] = S[p * 8 + m];
        interp_prims(sn, s, prims);
        double f, l[4];
        l[0] = 1.0;
        g.fl(s, f, l[1], l[2], l[3]);
        FluidScalars fs = fluid_frame(f, l, s, prims, cos_fallback);
        out[p] = fs.dens;
        out[n + p] = fs.u;
        out[2 * n + p] = acos(fs.cos_pitch);
        out[3 * n + p] = fs.kdotu;
        out[4 * n + p] = fs.b;
    }
}

__global__ void sample_prims_kernel(SnapshotView sn, const double* __restrict__ S, long n, double* __restrict__ out)
{
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        double x[4], prims[8];
#pragma unroll
        for (int m = 0; m < 4; m++) x[m] = S[p * 8 + m];
        interp_prims(sn, x, prims);
#pragma unroll
        for (int q = 0; q < 8; q++) out[q * n + p] = prims[q];
    }
}

static unsigned grid_for(long n, int threads)
{
    long blocks = (n + threads - 1) / threads;
    long cap = (long)sm_count() * 16;
    return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace mk
using namespace mk;

extern "C" int mk_snapshot_create(long nmb, long nk, long nj, long ni, const double* meshblocks,
                                  const int* prim_index, const double* geom, const int* grid, const int* gn,
                                  const double* g0, const double* ginv, const double* bbox_lo,
                                  const double* bbox_hi, int store_f32, mk_snapshot** out, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MK_REQUIRE(out != nullptr, "out is null");
    MK_REQUIRE(nmb > 0 && nk > 0 && nj > 0 && ni > 0, "empty snapshot");
    MK_REQUIRE(prim_index && geom && bbox_lo && bbox_hi, "null pointer");
    MK_REQUIRE(nmb < (1L << 31) && nk < 32768 && nj < 32768 && ni < 32768, "snapshot dimensions too large");
    PrimIndex pi;
    for (int q = 0; q < 8; q++) {
        MK_REQUIRE(prim_index[q] >= 0 && prim_index[q] < 8, "primitive index out of range");
        pi.p[q] = prim_index[q];
    }
    mk_snapshot* s = new mk_snapshot();
    memset(s, 0, sizeof *s);
    cudaGetDevice(&s->device);
    long cpb = (nk + 2) * (nj + 2) * (ni + 2);
    s->cell_bytes = nmb * cpb * 8 * (store_f32 ? 4 : 8);
    long geom_bytes = 12 * nmb * (long)sizeof(double);
    long grid_bytes = grid ? (long)gn[0] * gn[1] * gn[2] * (long)sizeof(int) : 0;
    s->total_bytes = s->cell_bytes + geom_bytes + grid_bytes;
    cudaError_t e = cudaMalloc(&s->cells, s->cell_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->geom, geom_bytes);
    if (e == cudaSuccess && grid) e = cudaMalloc((void**)&s->grid, grid_bytes);
    if (e != cudaSuccess) {
        set_error("snapshot allocation of %ld bytes failed: %s", s->total_bytes, cudaGetErrorString(e));
        mk_snapshot_destroy(s);
        return 1;
    }
    cudaMemcpyAsync(s->geom, geom, geom_bytes, cudaMemcpyDeviceToDevice, stream);
    if (grid) cudaMemcpyAsync(s->grid, grid, grid_bytes, cudaMemcpyDeviceToDevice, stream);
    if (!meshblocks) {
        // cells left uninitialised: the caller fills them (NCCL broadcast of a replicated snapshot)
    } else if (store_f32)
        repack_kernel<float><<<grid_for(nmb * cpb, 256), 256, 0, stream>>>(meshblocks, (float*)s->cells, nmb, cpb, pi);
    else
        repack_kernel<double><<<grid_for(nmb * cpb, 256), 256, 0, stream>>>(meshblocks, (double*)s->cells, nmb, cpb, pi);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_error("snapshot repack failed: %s", cudaGetErrorString(e));
        mk_snapshot_destroy(s);
        return 1;
    }
    SnapshotView& v = s->view;
    v.source = 0;
    v.cells = s->cells; v.is_f32 = store_f32 ? 1 : 0;
    v.nmb = (int)nmb; v.nk = (int)nk; v.nj = (int)nj; v.ni = (int)ni;
    v.geom = s->geom;
    for (int d = 0; d < 3; d++) {
        v.bbox_lo[d] = bbox_lo[d]; v.bbox_hi[d] = bbox_hi[d];
        v.gn[d] = grid ? gn[d] : 0;
        v.g0[d] = grid ? g0[d] : 0.0;
        v.ginv[d] = grid ? ginv[d] : 0.0;
    }
    v.grid = s->grid;
    *out = s;
    return 0;
}

extern "C" int mk_snapshot_create_torus(const double* params9, mk_snapshot** out)
{
    MK_REQUIRE(params9 && out, "null pointer");
    mk_snapshot* s = new mk_snapshot();
    memset(s, 0, sizeof *s);
    cudaGetDevice(&s->device);
    s->view.source = 1;
    TorusParams& t = s->view.torus;
    t.fluid_gamma = params9[0]; t.R0 = params9[1]; t.R_in = params9[2]; t.p = params9[3]; t.h = params9[4];
    t.u0 = params9[5]; t.beta0 = params9[6]; t.dens_scale = params9[7]; t.r_out = params9[8];
    *out = s;
    return 0;
}

extern "C" int mk_snapshot_destroy(mk_snapshot* s)
{
    if (!s) return 0;
    if (s->cells) cudaFree(s->cells);
    if (s->geom) cudaFree(s->geom);
    if (s->grid) cudaFree(s->grid);
    delete s;
    return 0;
}

extern "C" long mk_snapshot_bytes(const mk_snapshot* s) { return s ? s->total_bytes : 0; }

extern "C" int mk_snapshot_cells(mk_snapshot* s, void** cells, long* bytes)
{
    MK_REQUIRE(s != nullptr, "snapshot is null");
    if (cells) *cells = s->cells;
    if (bytes) *bytes = s->cell_bytes;
    return 0;
}

extern "C" int mk_sample_scalars(const mk_snapshot* snap, double bhspin, const double* S, long n,
                                 double fallback_pitch_angle, double* out, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(snap && S && out, "null pointer");
    KerrSchild g; g.a = bhspin; g.aa = bhspin * bhspin; g.rH = 0;
    sample_scalars_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(snap->view, g, S, n, cos(fallback_pitch_angle), out);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_sample_prims(const mk_snapshot* snap, const double* S, long n, double* out, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(snap && S && out, "null pointer");
    sample_prims_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(snap->view, S, n, out);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}
