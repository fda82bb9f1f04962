// Snapshot repack + stand-alone sampling kernels (the API-parity path of
// /root/reference/mahakala/grmhd/athenak.py:527-812; the fused path lives in render.cu).
#include <algorithm>
#include <utility>
#include <vector>
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

// ---------------------------------------------------------------------------------------------------------
// Ghost-zone fill fused with the repack (replaces the host loader loops of athenak.py:105-158 same-level copy
// :208-229 and the refinement-boundary branches :231-514).  Input: the interior arrays exactly as an .athdf
// file holds them, uov (nu, nmb, nk, nj, ni) and B (nb, nmb, nk, nj, ni), nu + nb = 8, float32 or float64.
// One thread per padded cell (mb, k, j, i): an interior cell copies its own values; a ghost cell in direction
// d = (di, dj, dk) takes
//   the edge cell of the same-level neighbour at LogicalLocation + d if that block exists, else
//   (refined meshes) the coarse cell that contains it (injection) if the coarser neighbour exists, else
//   the mean of the 8 finer cells it covers (summed in the reference's order: i offset outermost, k innermost,
//   then / 8; athenak.py:339-349, 405-422, 497-512) if all 8 exist,
//   else zero (domain boundary).
// Blocks are found by binary search in a table of (level, lk, lj, li) keys sorted on the host.
// ---------------------------------------------------------------------------------------------------------
struct BlockTable {
    const long long* keys;     // sorted
    const int* mb;             // block index of each key
    int n;
};

__host__ __device__ inline long long block_key(int lev, long li, long lj, long lk)
{
    return ((long long)lev << 57) | ((long long)lk << 38) | ((long long)lj << 19) | (long long)li;
}

__device__ __forceinline__ int find_block(const BlockTable& t, int lev, long li, long lj, long lk)
{
    const long lim = 1L << 19;
    if (lev < 0 || lev > 63 || li < 0 || lj < 0 || lk < 0 || li >= lim || lj >= lim || lk >= lim) return -1;
    long long key = block_key(lev, li, lj, lk);
    int lo = 0, hi = t.n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        long long v = t.keys[mid];
        if (v == key) return t.mb[mid];
        if (v < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

struct InteriorSrc {
    const void* uov;
    const void* B;
    int nu;                    // primitives held by uov; file index q >= nu lives in B[q - nu]
    long nmb, nk, nj, ni;
};

template <class SrcT>
__device__ __forceinline__ double interior_value(const InteriorSrc& s, int q, long mb, long k, long j, long i)
{
    const SrcT* base = (q < s.nu) ? reinterpret_cast<const SrcT*>(s.uov) + (long)q * s.nmb * s.nk * s.nj * s.ni
                                  : reinterpret_cast<const SrcT*>(s.B) + (long)(q - s.nu) * s.nmb * s.nk * s.nj * s.ni;
    return (double)base[((mb * s.nk + k) * s.nj + j) * s.ni + i];
}

__device__ __forceinline__ long floor_div(long a, long b) { long q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

template <class SrcT, class CellT>
__global__ void ghost_fill_repack_kernel(InteriorSrc src, BlockTable tab, const int* __restrict__ loc,
                                         const int* __restrict__ lev, int multilevel, PrimIndex pi,
                                         CellT* __restrict__ dst, int* __restrict__ lossy)
{
    const long pk = src.nk + 2, pj = src.nj + 2, pi_ = src.ni + 2;
    const long cpb = pk * pj * pi_, total = src.nmb * cpb;
    const long n[3] = {src.ni, src.nj, src.nk};
    bool any_lossy = false;
    for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < total; c += (long)gridDim.x * blockDim.x) {
        long mb = c / cpb, w = c - mb * cpb;
        long t[3];                                   // padded index along (i, j, k)
        t[2] = w / (pj * pi_); w -= t[2] * pj * pi_;
        t[1] = w / pi_; t[0] = w - t[1] * pi_;
        int d[3];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) d[ax] = (t[ax] == 0) ? -1 : ((t[ax] == n[ax] + 1) ? 1 : 0);
        double v[8];                                 // file order
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = 0.0;
        if ((d[0] | d[1] | d[2]) == 0) {
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = interior_value<SrcT>(src, q, mb, t[2] - 1, t[1] - 1, t[0] - 1);
        } else {
            const int L = lev[mb];
            const long l[3] = {loc[3 * mb], loc[3 * mb + 1], loc[3 * mb + 2]};
            int nb = find_block(tab, L, l[0] + d[0], l[1] + d[1], l[2] + d[2]);
            if (nb >= 0) {                           // athenak.py:208-229: same-level neighbour's edge cell
                long s_[3];
#pragma unroll
                for (int ax = 0; ax < 3; ax++) s_[ax] = (d[ax] == 1) ? 0 : ((d[ax] == -1) ? n[ax] - 1 : t[ax] - 1);
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = interior_value<SrcT>(src, q, nb, s_[2], s_[1], s_[0]);
            } else if (multilevel) {
                long g[3], cl[3];                    // global cell index at this block's level; coarse block
#pragma unroll
                for (int ax = 0; ax < 3; ax++) { g[ax] = l[ax] * n[ax] + (t[ax] - 1); cl[ax] = floor_div(l[ax] + d[ax], 2); }
                int cb = find_block(tab, L - 1, cl[0], cl[1], cl[2]);
                if (cb >= 0) {                       // coarser neighbour: injection
                    long s_[3];
                    bool ok = true;
#pragma unroll
                    for (int ax = 0; ax < 3; ax++) {
                        s_[ax] = floor_div(g[ax], 2) - cl[ax] * n[ax];
                        ok &= (s_[ax] >= 0) & (s_[ax] < n[ax]);
                    }
                    if (ok) {
#pragma unroll
                        for (int q = 0; q < 8; q++) v[q] = interior_value<SrcT>(src, q, cb, s_[2], s_[1], s_[0]);
                    }
                } else {                             // finer neighbours: mean of the 8 covered cells
                    double acc[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = 0.0;
                    int cnt = 0;
                    for (int o = 0; o < 8; o++) {    // o = oi*4 + oj*2 + ok: the reference's summation order
                        long f[3] = {2 * g[0] + ((o >> 2) & 1), 2 * g[1] + ((o >> 1) & 1), 2 * g[2] + (o & 1)};
                        long fb[3], fc[3];
#pragma unroll
                        for (int ax = 0; ax < 3; ax++) { fb[ax] = floor_div(f[ax], n[ax]); fc[ax] = f[ax] - fb[ax] * n[ax]; }
                        int fm = find_block(tab, L + 1, fb[0], fb[1], fb[2]);
                        if (fm < 0) continue;
                        cnt++;
#pragma unroll
                        for (int q = 0; q < 8; q++) acc[q] += interior_value<SrcT>(src, q, fm, fc[2], fc[1], fc[0]);
                    }
                    if (cnt == 8) {
#pragma unroll
                        for (int q = 0; q < 8; q++) v[q] = acc[q] / 8.0;
                    }
                }
            }
        }
        CellT o8[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            double x = v[pi.p[q]];
            o8[q] = (CellT)x;
            if (sizeof(CellT) == 4) any_lossy |= !((double)o8[q] == x);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) dst[c * 8 + q] = o8[q];
    }
    if (lossy && any_lossy) atomicOr(lossy, 1);
}

// cells[mb][k][j][i][8] (canonical order) -> the reference's all_meshblocks (nmb, 8, nk+2, nj+2, ni+2) in file order
template <class CellT>
__global__ void unpack_kernel(const CellT* __restrict__ cells, double* __restrict__ out, long nmb, long cells_per_block,
                              PrimIndex pi)
{
    long total = nmb * cells_per_block;
    for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < total; c += (long)gridDim.x * blockDim.x) {
        long mb = c / cells_per_block, w = c - mb * cells_per_block;
#pragma unroll
        for (int q = 0; q < 8; q++) out[(mb * 8 + pi.p[q]) * cells_per_block + w] = (double)cells[c * 8 + q];
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

// caller's (nmb, 12) geometry records -> internal (nmb, 16) records with 1/dx; *not_pow2 is raised when some cell
// size is not a power of two (then 1/dx is inexact and cell_index keeps the corrected division)
__global__ void geom_expand_kernel(const double* __restrict__ g12, double* __restrict__ g16, long nmb, int* not_pow2)
{
    long mb = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (mb >= nmb) return;
    const double* s = g12 + mb * 12;
    double* d = g16 + mb * GEOM_DOUBLES;
    for (int i = 0; i < 12; i++) d[i] = s[i];
    bool bad = false;
    for (int a = 0; a < 3; a++) {
        double dx = s[9 + a];
        d[12 + a] = 1.0 / dx;
        unsigned long long bits = (unsigned long long)__double_as_longlong(dx);
        unsigned expo = (unsigned)((bits >> 52) & 0x7ffu);
        bad |= !(dx > 0.0) || (bits & ((1ULL << 52) - 1ULL)) != 0 || expo < 64 || expo > 1983;
    }
    d[15] = 0.0;
    if (bad) atomicOr(not_pow2, 1);
}

static unsigned grid_for(long n, int threads)
{
    long blocks = (n + threads - 1) / threads;
    long cap = (long)sm_count() * 16;
    return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace mk
using namespace mk;

// allocation + view set-up shared by the two snapshot constructors; *out receives the handle (cells unfilled)
static int snapshot_alloc(long nmb, long nk, long nj, long ni, const double* geom, const int* grid, const int* gn,
                          const double* g0, const double* ginv, const double* bbox_lo, const double* bbox_hi,
                          int store_f32, mk_snapshot** out, cudaStream_t stream)
{
    MK_REQUIRE(out != nullptr, "out is null");
    MK_REQUIRE(nmb > 0 && nk > 0 && nj > 0 && ni > 0, "empty snapshot");
    MK_REQUIRE(geom && bbox_lo && bbox_hi, "null pointer");
    MK_REQUIRE(!grid || (gn && g0 && ginv), "grid given without gn / g0 / ginv");
    MK_REQUIRE(nmb < (1L << 31) && nk < 32768 && nj < 32768 && ni < 32768, "snapshot dimensions too large");
    mk_snapshot* s = new mk_snapshot();
    memset(s, 0, sizeof *s);
    cudaGetDevice(&s->device);
    long cpb = (nk + 2) * (nj + 2) * (ni + 2);
    s->cell_bytes = nmb * cpb * 8 * (store_f32 ? 4 : 8);
    long geom_bytes = GEOM_DOUBLES * nmb * (long)sizeof(double);
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
    int* flag = (int*)queue_counter(stream, 2);       // zeroed device word
    if (!flag) { mk_snapshot_destroy(s); return 1; }
    geom_expand_kernel<<<(unsigned)((nmb + 127) / 128), 128, 0, stream>>>(geom, s->geom, nmb, flag);
    int not_pow2 = 1;
    e = cudaMemcpyAsync(&not_pow2, flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_error("snapshot geometry set-up failed: %s", cudaGetErrorString(e));
        mk_snapshot_destroy(s);
        return 1;
    }
    if (grid) cudaMemcpyAsync(s->grid, grid, grid_bytes, cudaMemcpyDeviceToDevice, stream);
    SnapshotView& v = s->view;
    v.dx_pow2 = not_pow2 ? 0 : 1;
    v.source = 0;
    v.cells = s->cells; v.is_f32 = store_f32 ? 1 : 0;
    v.nmb = (int)nmb; v.nk = (int)nk; v.nj = (int)nj; v.ni = (int)ni;
    v.sj = (ni + 2) * 8; v.sk = v.sj * (nj + 2); v.sb = v.sk * (nk + 2);
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

static int read_prim_index(const int* prim_index, PrimIndex& pi)
{
    MK_REQUIRE(prim_index != nullptr, "prim_index is null");
    int seen = 0;
    for (int q = 0; q < 8; q++) {
        MK_REQUIRE(prim_index[q] >= 0 && prim_index[q] < 8, "primitive index out of range");
        pi.p[q] = prim_index[q];
        seen |= 1 << prim_index[q];
    }
    MK_REQUIRE(seen == 0xff, "prim_index must be a permutation of 0..7");
    return 0;
}

extern "C" int mk_snapshot_create(long nmb, long nk, long nj, long ni, const double* meshblocks,
                                  const int* prim_index, const double* geom, const int* grid, const int* gn,
                                  const double* g0, const double* ginv, const double* bbox_lo,
                                  const double* bbox_hi, int store_f32, mk_snapshot** out, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PrimIndex pi;
    if (int rc = read_prim_index(prim_index, pi)) return rc;
    mk_snapshot* s = nullptr;
    if (int rc = snapshot_alloc(nmb, nk, nj, ni, geom, grid, gn, g0, ginv, bbox_lo, bbox_hi, store_f32, &s, stream))
        return rc;
    long cpb = (nk + 2) * (nj + 2) * (ni + 2);
    if (!meshblocks) {
        // cells left uninitialised: the caller fills them (NCCL broadcast of a replicated snapshot)
    } else if (store_f32)
        repack_kernel<float><<<grid_for(nmb * cpb, 256), 256, 0, stream>>>(meshblocks, (float*)s->cells, nmb, cpb, pi);
    else
        repack_kernel<double><<<grid_for(nmb * cpb, 256), 256, 0, stream>>>(meshblocks, (double*)s->cells, nmb, cpb, pi);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_error("snapshot repack failed: %s", cudaGetErrorString(e));
        mk_snapshot_destroy(s);
        return 1;
    }
    *out = s;
    return 0;
}

extern "C" int mk_snapshot_create_from_interiors(long nmb, long nk, long nj, long ni, const void* uov, int n_uov,
                                                 const void* B, int n_B, int src_f32, const int* prim_index,
                                                 const int* logical_locations, const int* levels,
                                                 const double* geom, const int* grid, const int* gn,
                                                 const double* g0, const double* ginv, const double* bbox_lo,
                                                 const double* bbox_hi, int store_mode, int* stored_f32,
                                                 mk_snapshot** out, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MK_REQUIRE(uov && B && logical_locations && levels, "null pointer");
    MK_REQUIRE(n_uov >= 0 && n_B >= 0 && n_uov + n_B == 8, "uov and B must hold 8 primitives together");
    MK_REQUIRE(store_mode >= 0 && store_mode <= 2, "store_mode must be 0 (f64), 1 (f32) or 2 (auto)");
    MK_REQUIRE(nmb > 0 && nmb < (1L << 31), "bad block count");
    PrimIndex pi;
    if (int rc = read_prim_index(prim_index, pi)) return rc;
    // (level, location) -> block table, sorted on the host (athenak.py:160-206 builds a dict for the same purpose)
    std::vector<std::pair<long long, int>> tab((size_t)nmb);
    int lev_min = levels[0], lev_max = levels[0];
    for (long mb = 0; mb < nmb; mb++) {
        const int* l = logical_locations + 3 * mb;
        MK_REQUIRE(levels[mb] >= 0 && levels[mb] < 64, "refinement level out of range");
        MK_REQUIRE(l[0] >= 0 && l[1] >= 0 && l[2] >= 0 && l[0] < (1 << 19) && l[1] < (1 << 19) && l[2] < (1 << 19),
                   "logical location out of range");
        tab[(size_t)mb] = {block_key(levels[mb], l[0], l[1], l[2]), (int)mb};
        lev_min = levels[mb] < lev_min ? levels[mb] : lev_min;
        lev_max = levels[mb] > lev_max ? levels[mb] : lev_max;
    }
    std::sort(tab.begin(), tab.end());
    for (long i = 1; i < nmb; i++)
        MK_REQUIRE(tab[(size_t)i].first != tab[(size_t)i - 1].first, "two meshblocks share (level, logical location)");
    std::vector<long long> keys((size_t)nmb);
    std::vector<int> kmb((size_t)nmb);
    for (long i = 0; i < nmb; i++) { keys[(size_t)i] = tab[(size_t)i].first; kmb[(size_t)i] = tab[(size_t)i].second; }
    // one scratch allocation: keys | block of key | locations | levels | lossy flag
    size_t off_mb = (size_t)nmb * 8, off_loc = off_mb + (size_t)nmb * 4, off_lev = off_loc + (size_t)nmb * 12,
           off_flag = off_lev + (size_t)nmb * 4, scratch_bytes = off_flag + 8;
    char* scratch = nullptr;
    MK_CUDA_CHECK(cudaMalloc((void**)&scratch, scratch_bytes));
    cudaMemcpyAsync(scratch, keys.data(), (size_t)nmb * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(scratch + off_mb, kmb.data(), (size_t)nmb * 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(scratch + off_loc, logical_locations, (size_t)nmb * 12, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(scratch + off_lev, levels, (size_t)nmb * 4, cudaMemcpyHostToDevice, stream);
    cudaMemsetAsync(scratch + off_flag, 0, 8, stream);
    BlockTable bt;
    bt.keys = (const long long*)scratch; bt.mb = (const int*)(scratch + off_mb); bt.n = (int)nmb;
    InteriorSrc src;
    src.uov = uov; src.B = B; src.nu = n_uov; src.nmb = nmb; src.nk = nk; src.nj = nj; src.ni = ni;
    const int* d_loc = (const int*)(scratch + off_loc);
    const int* d_lev = (const int*)(scratch + off_lev);
    int* d_flag = (int*)(scratch + off_flag);
    const int multilevel = lev_max > lev_min;
    const long cpb = (nk + 2) * (nj + 2) * (ni + 2);
    int rc = 0;
    mk_snapshot* s = nullptr;
    // auto: try float32 cells; the kernel raises a flag if any stored value (ghost averages included) does not
    // survive the round trip, in which case the snapshot is rebuilt with float64 cells
    for (int attempt = 0; attempt < 2 && rc == 0; attempt++) {
        const int f32 = (store_mode == 2) ? (attempt == 0) : store_mode;
        rc = snapshot_alloc(nmb, nk, nj, ni, geom, grid, gn, g0, ginv, bbox_lo, bbox_hi, f32, &s, stream);
        if (rc) break;
        const unsigned g = grid_for(nmb * cpb, 256);
        if (f32 && src_f32)
            ghost_fill_repack_kernel<float, float><<<g, 256, 0, stream>>>(src, bt, d_loc, d_lev, multilevel, pi, (float*)s->cells, d_flag);
        else if (f32)
            ghost_fill_repack_kernel<double, float><<<g, 256, 0, stream>>>(src, bt, d_loc, d_lev, multilevel, pi, (float*)s->cells, d_flag);
        else if (src_f32)
            ghost_fill_repack_kernel<float, double><<<g, 256, 0, stream>>>(src, bt, d_loc, d_lev, multilevel, pi, (double*)s->cells, nullptr);
        else
            ghost_fill_repack_kernel<double, double><<<g, 256, 0, stream>>>(src, bt, d_loc, d_lev, multilevel, pi, (double*)s->cells, nullptr);
        int flag = 0;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof flag, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            set_error("ghost fill / repack failed: %s", cudaGetErrorString(e));
            rc = 1;
        } else if (store_mode == 2 && f32 && flag) {
            mk_snapshot_destroy(s);          // lossy as float32: second attempt stores float64
            s = nullptr;
            continue;
        } else {
            if (stored_f32) *stored_f32 = f32;
        }
        break;
    }
    cudaFree(scratch);
    if (rc) {
        if (s) mk_snapshot_destroy(s);
        return rc;
    }
    *out = s;
    return 0;
}

extern "C" int mk_snapshot_unpack(const mk_snapshot* snap, const int* prim_index, double* meshblocks, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MK_REQUIRE(snap && meshblocks, "null pointer");
    MK_REQUIRE(snap->view.source == 0 && snap->cells, "snapshot holds no cells");
    PrimIndex pi;
    if (int rc = read_prim_index(prim_index, pi)) return rc;
    const SnapshotView& v = snap->view;
    long cpb = (long)(v.nk + 2) * (v.nj + 2) * (v.ni + 2);
    if (v.is_f32)
        unpack_kernel<float><<<grid_for(v.nmb * cpb, 256), 256, 0, stream>>>((const float*)snap->cells, meshblocks, v.nmb, cpb, pi);
    else
        unpack_kernel<double><<<grid_for(v.nmb * cpb, 256), 256, 0, stream>>>((const double*)snap->cells, meshblocks, v.nmb, cpb, pi);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_snapshot_create_torus(const double* params9, mk_snapshot** out)
{
    MK_REQUIRE(params9 && out, "null pointer");
    MK_REQUIRE(params9[1] > 0.0 && params9[4] > 0.0 && params9[6] > 0.0, "torus parameters R0, h, beta0 must be positive");
    mk_snapshot* s = new mk_snapshot();
    memset(s, 0, sizeof *s);
    cudaGetDevice(&s->device);
    s->view.source = 1;
    TorusParams& t = s->view.torus;
    t.fluid_gamma = params9[0]; t.R0 = params9[1]; t.R_in = params9[2]; t.p = params9[3]; t.h = params9[4];
    t.u0 = params9[5]; t.beta0 = params9[6]; t.dens_scale = params9[7]; t.r_out = params9[8];
    t.inv_R0 = 1.0 / t.R0; t.inv_2h2 = 1.0 / (2.0 * t.h * t.h); t.cB = 2.0 * (t.fluid_gamma - 1.0) / t.beta0;
    t.u0R0 = t.u0 * t.R0;
    t.p_is_three_halves = (t.p == 1.5);
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
    KerrSchild g; g.set_spin(bhspin);
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
