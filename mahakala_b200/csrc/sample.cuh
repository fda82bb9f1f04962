// GRMHD-snapshot sampling and emission physics as device functions (used by the stand-alone sampling
// / synchrotron kernels and by the fused render kernel).
//
// Restates /root/reference/mahakala/
//   grmhd/athenak.py:663-670   point -> meshblock (left-open / right-closed extents)
//   grmhd/athenak.py:718-757   ghost-padded trilinear interpolation of the 8 primitives
//   grmhd/athenak.py:760-794   u^m, b^m, k.u, k.b, b.b, pitch angle
//   images.py:87-118           beta, sigma, units, local frequency, sigma cut
//   electrons.py:46-50         R_low / R_high electron temperature
//   transfer.py:56-86          thermal synchrotron j_nu, alpha_nu (invariant form)
//
// HBM layout (repacked once at snapshot creation; the reference re-uploads (nmb, 8, nk+2, nj+2, ni+2)
// on every call, athenak.py:693):  cells[mb][k][j][i][8]  with the 8 primitives of one cell contiguous
// (64 B as f64, 32 B as f32) in canonical order  dens, eint, U1, U2, U3, B1, B2, B3.  The two x-adjacent
// corners of a sample are one 128 B (f64) line; a sample touches 4 such segments instead of 64 sectors.
#pragma once
#include "fp64_math.cuh"

namespace mk {

// cfg3 analytic thin torus: primitives are closed-form functions of position.
struct TorusParams {
    double fluid_gamma, R0, R_in, p, h, u0, beta0, dens_scale, r_out;
    // derived on the host (mk_snapshot_create_torus)
    double inv_R0, inv_2h2, cB, u0R0;     // 1/R0, 1/(2 h^2), 2 (gamma - 1)/beta0, u0 R0
    int p_is_three_halves;                // the default power law: x^-1.5 = rsqrt(x)^3 instead of pow()
};

struct SnapshotView {
    int source;               // 0 = AthenaK-style snapshot cells, 1 = analytic torus
    TorusParams torus;
    const void* cells;        // AoS cells, f64 or f32
    int is_f32;
    int nmb, nk, nj, ni;      // interior cells per block
    long sj, sk, sb;          // element strides of the padded cell array: next j row, next k plane, next block
    // per-block geometry (nmb, 16): lo[3], hi[3] face extents, v0[3] first cell centre, dx[3] cell size, 1/dx[3],
    // pad; one 128 B record per block so that a lookup costs one round of 256-bit loads instead of dependent ones
    const double* geom;
    // every cell size of the mesh is a power of two (AthenaK meshes on power-of-two domains): 1/dx is exact, so
    // xi * (1/dx) IS xi / dx and the floor division of athenak.py:718-733 needs neither a division nor a fix-up
    int dx_pow2;
    // block lookup grid over the bounding box (regular meshes); grid == nullptr -> linear scan
    const int* grid;
    int gn[3];
    double g0[3], ginv[3];
    double bbox_lo[3], bbox_hi[3];
};

constexpr int GEOM_DOUBLES = 16;

struct BlockGeom {
    double lo[3], hi[3], v0[3], dx[3];
};

MK_HD BlockGeom load_geom(const SnapshotView& sn, int mb)
{
    const double* p = sn.geom + (long)mb * GEOM_DOUBLES;
    double v[12];
#if defined(__CUDA_ARCH__) && !defined(__CUDACC_RTC__)
#pragma unroll
    for (int i = 0; i < 3; i++)
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
            : "=d"(v[4 * i]), "=d"(v[4 * i + 1]), "=d"(v[4 * i + 2]), "=d"(v[4 * i + 3]) : "l"(p + 4 * i));
#elif defined(__CUDA_ARCH__)    // NVRTC 12.9's embedded ptxas rejects 256-bit vector accesses: 128-bit loads
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double2 t = __ldg(reinterpret_cast<const double2*>(p) + i);
        v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
#else
    for (int i = 0; i < 12; i++) v[i] = p[i];
#endif
    BlockGeom g;
#pragma unroll
    for (int d = 0; d < 3; d++) { g.lo[d] = v[d]; g.hi[d] = v[3 + d]; g.v0[d] = v[6 + d]; g.dx[d] = v[9 + d]; }
    return g;
}

// 1/dx of the block: the last 32 B of its 128 B record (same line as load_geom just touched), fetched where it is
// needed so that it does not lengthen the live range of the record
MK_HD void load_inv_dx(const SnapshotView& sn, int mb, double inv[3])
{
    const double* p = sn.geom + (long)mb * GEOM_DOUBLES + 12;
#if defined(__CUDA_ARCH__) && !defined(__CUDACC_RTC__)
    double pad;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(inv[0]), "=d"(inv[1]), "=d"(inv[2]), "=d"(pad) : "l"(p));
    (void)pad;
#elif defined(__CUDA_ARCH__)
    double2 t0 = __ldg(reinterpret_cast<const double2*>(p)), t1 = __ldg(reinterpret_cast<const double2*>(p) + 1);
    inv[0] = t0.x; inv[1] = t0.y; inv[2] = t1.x;
#else
    for (int i = 0; i < 3; i++) inv[i] = p[i];
#endif
}

// left-open / right-closed membership (athenak.py:666-668), branch-free
MK_HD bool in_block(const BlockGeom& g, const double x[4])
{
    return (g.lo[0] < x[1]) & (x[1] <= g.hi[0]) & (g.lo[1] < x[2]) & (x[2] <= g.hi[1]) & (g.lo[2] < x[3]) &
           (x[3] <= g.hi[2]);
}

// Rare continuation of the grid lookup, kept OUT OF LINE so that the fused kernel's hot loop does not carry its code:
// a point within rounding distance of a face may land in the neighbouring grid cell -> fix up with the exact extents
// (x <= lo -> step down, x > hi -> step up), at most one step per axis; then, as a last resort (holes in the mesh,
// degenerate geometry), the exhaustive 3x3x3 neighbourhood.
// (All arguments by value: taking the address of the kernel-parameter SnapshotView would force a stack copy of it.)
#ifdef __CUDA_ARCH__
static __device__ __noinline__
#else
static inline
#endif
int locate_block_slow(const int* grid, const double* geom, int gn0, int gn1, int gn2, double x1, double x2, double x3,
                      int c0, int c1, int c2, int mb)
{
    if (!((x1 == x1) & (x2 == x2) & (x3 == x3))) return -1;          // NaN position: outside
    SnapshotView sn;
    sn.geom = geom;
    const double x[4] = {0.0, x1, x2, x3};
    const int c[3] = {c0, c1, c2}, gn[3] = {gn0, gn1, gn2};
    int c2v[3];
    BlockGeom geo;
    if (mb >= 0) geo = load_geom(sn, mb);
    for (int d = 0; d < 3; d++) {
        int s = 0;
        if (mb >= 0) s = (x[d + 1] <= geo.lo[d]) ? -1 : ((x[d + 1] > geo.hi[d]) ? 1 : 0);
        c2v[d] = min(max(c[d] + s, 0), gn[d] - 1);
    }
    int mb2 = grid[(c2v[2] * gn1 + c2v[1]) * gn0 + c2v[0]];
    if (mb2 >= 0) {
        geo = load_geom(sn, mb2);
        if (in_block(geo, x)) return mb2;
    }
    for (int dk = -1; dk <= 1; dk++)
        for (int dj = -1; dj <= 1; dj++)
            for (int di = -1; di <= 1; di++) {
                int a = c0 + di, b = c1 + dj, e = c2 + dk;
                if (a < 0 || b < 0 || e < 0 || a >= gn0 || b >= gn1 || e >= gn2) continue;
                int m = grid[(e * gn1 + b) * gn0 + a];
                if (m < 0) continue;
                geo = load_geom(sn, m);
                if (in_block(geo, x)) return m;
            }
    return -1;
}

// Grid lookup (regular meshes): O(1) replacement of the reference's mask loop over meshblocks, athenak.py:663-670.
// Grid cell of the point.  The range test on the integer cell index doubles as the bounding-box rejection:
// x below the lower domain face gives a negative quotient (x - g0 is negative exactly when x < g0), x above
// the upper face gives q >= gn.  q == gn is the one ambiguous value (x on the upper face, which is inside --
// right-closed extents -- or just beyond it), so only there the exact face is consulted.  NaN converts to 0
// and fails the exact membership test below.  Whatever passes is decided by in_block() on the stored faces.
MK_HD int locate_block_grid(const SnapshotView& sn, const double x[4], BlockGeom& geo)
{
    int c[3];
    bool maybe = true;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        int q = floor_to_int((x[d + 1] - sn.g0[d]) * sn.ginv[d]);
        if (q == sn.gn[d]) q = (x[d + 1] <= sn.bbox_hi[d]) ? q - 1 : q;
        maybe &= ((unsigned)q < (unsigned)sn.gn[d]);
        c[d] = q;
    }
    if (!maybe) return -1;
    int mb = sn.grid[(c[2] * sn.gn[1] + c[1]) * sn.gn[0] + c[0]];
    if (mb >= 0) {
        geo = load_geom(sn, mb);
        if (in_block(geo, x)) return mb;
    }
    mb = locate_block_slow(sn.grid, sn.geom, sn.gn[0], sn.gn[1], sn.gn[2], x[1], x[2], x[3], c[0], c[1], c[2], mb);
    if (mb >= 0) geo = load_geom(sn, mb);
    return mb;
}

// athenak.py:663-670.  Blocks tile the domain without overlap, so "last match wins" == "the match".
// Returns the block index (or -1) and its geometry record.
MK_HD int locate_block(const SnapshotView& sn, const double x[4], BlockGeom& geo)
{
    if (sn.grid == nullptr) {
        // NaN-safe bounding-box rejection (comparisons with NaN are false -> outside), then the reference's scan
        if (!((sn.bbox_lo[0] < x[1]) & (x[1] <= sn.bbox_hi[0]) & (sn.bbox_lo[1] < x[2]) & (x[2] <= sn.bbox_hi[1]) &
              (sn.bbox_lo[2] < x[3]) & (x[3] <= sn.bbox_hi[2])))
            return -1;
        int found = -1;
        for (int mb = 0; mb < sn.nmb; mb++) {
            BlockGeom g = load_geom(sn, mb);
            if (in_block(g, x)) { found = mb; geo = g; }
        }
        return found;
    }
    return locate_block_grid(sn, x, geo);
}

// athenak.py:718-733: xi = x - x_v[0] + dx ; idx = xi // dx ; delta = (xi / dx) % 1.
// pow2 (warp-uniform): dx is a power of two, so xi * (1/dx) is the exact quotient; its floor and fractional part are
// what the reference's floor division and "% 1" return, with no rounding to repair.
MK_HD void cell_index(double x, double v0, double dx, double inv_dx, bool pow2, int& idx, double& delta)
{
    double xi = x - v0 + dx;
    if (pow2) {
        double qd = xi * inv_dx;
        double q = floor(qd);
        delta = qd - q;
        idx = (int)q;
        return;
    }
    double qd = fast_div(xi, dx);         // dx is a positive normal number
    double q = floor(qd);
    delta = qd - q;                       // python float % 1. for qd >= 0
    double rem = fma(-q, dx, xi);         // exact floor division (np.floor_divide works from fmod)
    idx = (int)q;
    if (rem < 0.0) idx -= 1;
    if (rem >= dx) idx += 1;
}

MK_HD void load_cell_pair(const double* p, double a[8], double b[8])
{
    // two x-adjacent cells = 128 contiguous bytes, read as four 256-bit read-only loads
    double4 v[4];
#if defined(__CUDA_ARCH__) && !defined(__CUDACC_RTC__)
#pragma unroll
    for (int i = 0; i < 4; i++)
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
            : "=d"(v[i].x), "=d"(v[i].y), "=d"(v[i].z), "=d"(v[i].w) : "l"(p + 4 * i));
#elif defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double2 lo = __ldg(reinterpret_cast<const double2*>(p) + 2 * i), hi = __ldg(reinterpret_cast<const double2*>(p) + 2 * i + 1);
        v[i].x = lo.x; v[i].y = lo.y; v[i].z = hi.x; v[i].w = hi.y;
    }
#else
    for (int i = 0; i < 4; i++) v[i] = reinterpret_cast<const double4*>(p)[i];
#endif
    a[0] = v[0].x; a[1] = v[0].y; a[2] = v[0].z; a[3] = v[0].w; a[4] = v[1].x; a[5] = v[1].y; a[6] = v[1].z; a[7] = v[1].w;
    b[0] = v[2].x; b[1] = v[2].y; b[2] = v[2].z; b[3] = v[2].w; b[4] = v[3].x; b[5] = v[3].y; b[6] = v[3].z; b[7] = v[3].w;
}

MK_HD void load_cell_pair(const float* p, double a[8], double b[8])
{
    // two x-adjacent f32 cells = 64 contiguous bytes, two 256-bit read-only loads
    float v[16];
#if defined(__CUDA_ARCH__) && defined(__CUDACC_RTC__)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
#elif defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < 2; i++)
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(v[8 * i]), "=f"(v[8 * i + 1]), "=f"(v[8 * i + 2]), "=f"(v[8 * i + 3]), "=f"(v[8 * i + 4]),
              "=f"(v[8 * i + 5]), "=f"(v[8 * i + 6]), "=f"(v[8 * i + 7])
            : "l"(p + 8 * i));
#else
    for (int i = 0; i < 16; i++) v[i] = p[i];
#endif
#pragma unroll
    for (int q = 0; q < 8; q++) { a[q] = (double)v[q]; b[q] = (double)v[8 + q]; }
}

// athenak.py:737-752: 8 corners x 8 primitives.  The reference nests lerps in a + (b - a) t form; here the
// same trilinear interpolant is accumulated as sum_c w_c v_c with the 8 corner weights (identical in exact
// arithmetic, ~1 ulp apart in floating point): 64 FMAs instead of 112 operations, and each x-pair of cells is
// consumed as soon as it is loaded, which keeps the register footprint of the fused kernel small.
template <class CellT, int POW2 = -1>
MK_HD void trilinear(const SnapshotView& sn, int mb, const BlockGeom& geo, const double x[4],
                                          double prims[8])
{
    int i1, i2, i3;
    double d1, d2, d3;
    const bool pow2 = (POW2 < 0) ? (sn.dx_pow2 != 0) : (POW2 != 0);
    double inv[3] = {0.0, 0.0, 0.0};
    if (pow2) load_inv_dx(sn, mb, inv);
    cell_index(x[1], geo.v0[0], geo.dx[0], inv[0], pow2, i1, d1);
    cell_index(x[2], geo.v0[1], geo.dx[1], inv[1], pow2, i2, d2);
    cell_index(x[3], geo.v0[2], geo.dx[2], inv[2], pow2, i3, d3);
    // in-block points have indices in [0, n]; clamp defensively so that no load can leave the block
    i1 = min(max(i1, 0), sn.ni); i2 = min(max(i2, 0), sn.nj); i3 = min(max(i3, 0), sn.nk);
    const long sj = sn.sj, sk = sn.sk;
    const CellT* base = reinterpret_cast<const CellT*>(sn.cells) + (mb * sn.sb + i3 * sk + i2 * sj + (long)(i1 * 8));
    const double e1 = 1.0 - d1, e2 = 1.0 - d2, e3 = 1.0 - d3;
#pragma unroll
    for (int q = 0; q < 8; q++) prims[q] = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const double w = ((c & 2) ? d3 : e3) * ((c & 1) ? d2 : e2);
        const double w0 = w * e1, w1 = w * d1;
        double a[8], b[8];
        load_cell_pair(base + ((c & 2) ? sk : 0) + ((c & 1) ? sj : 0), a, b);
#pragma unroll
        for (int q = 0; q < 8; q++) prims[q] = fma(w0, a[q], fma(w1, b[q], prims[q]));
    }
}

#if defined(MK_RENDER_SMEM_STAGE) && !defined(__CUDACC_RTC__)
// EXPERIMENT (compiled only with -DMK_RENDER_SMEM_STAGE; scripts/build_variant.sh): the north_star's "hot cells staged in
// shared memory".  The lanes of a warp sample a 4x8-pixel patch, far smaller than a cell, so the union of their 2x2x2
// corner cells is almost always one 3x3x3 brick of a single meshblock.  One elected lane fetches that brick with nine
// TMA bulk copies (cp.async.bulk, 3 cells = 192 B per row) signalled on the warp's mbarrier; every lane then reads its
// eight corners from shared memory.  Measured against the direct 256-bit read-only loads: see DESIGN.md.
struct CellStage {
    double* brick;                 // [3][3][3][8] doubles of this warp
    unsigned long long* bar;       // the warp's mbarrier
};

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void trilinear_staged(const SnapshotView& sn, int mb, const BlockGeom& geo, const double x[4],
                                                 double prims[8], const CellStage& st)
{
    int i1, i2, i3;
    double d1, d2, d3;
    double inv[3];
    load_inv_dx(sn, mb, inv);
    cell_index(x[1], geo.v0[0], geo.dx[0], inv[0], true, i1, d1);
    cell_index(x[2], geo.v0[1], geo.dx[1], inv[1], true, i2, d2);
    cell_index(x[3], geo.v0[2], geo.dx[2], inv[2], true, i3, d3);
    i1 = min(max(i1, 0), sn.ni); i2 = min(max(i2, 0), sn.nj); i3 = min(max(i3, 0), sn.nk);
    const long sj = sn.sj, sk = sn.sk;
    const unsigned mask = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    const int mb0 = __shfl_sync(mask, mb, leader);
    const int m1 = __reduce_min_sync(mask, i1), m2 = __reduce_min_sync(mask, i2), m3 = __reduce_min_sync(mask, i3);
    const int M1 = __reduce_max_sync(mask, i1), M2 = __reduce_max_sync(mask, i2), M3 = __reduce_max_sync(mask, i3);
    // the brick starts at (m3, m2, m1) and spans 3 cells per axis; rows must stay inside the padded block
    const bool fits = __all_sync(mask, mb == mb0) && (M1 - m1 <= 1) && (M2 - m2 <= 1) && (M3 - m3 <= 1) &&
                      (m1 + 2 <= sn.ni + 1) && (m2 + 2 <= sn.nj + 1) && (m3 + 2 <= sn.nk + 1);
    const double e1 = 1.0 - d1, e2 = 1.0 - d2, e3 = 1.0 - d3;
#pragma unroll
    for (int q = 0; q < 8; q++) prims[q] = 0.0;
    if (fits) {
        unsigned long long token = 0;
        const unsigned bar = smem_addr(st.bar);
        if ((int)lane == leader) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic reads of the brick are done
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(token) : "r"(bar), "r"(27 * 64) : "memory");
            const double* base = reinterpret_cast<const double*>(sn.cells) + (mb0 * sn.sb + m3 * sk + m2 * sj + (long)(m1 * 8));
#pragma unroll
            for (int r = 0; r < 9; r++) {
                const double* src = base + (r / 3) * sk + (r % 3) * sj;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_addr(st.brick + r * 24)), "l"(src), "r"(192), "r"(bar) : "memory");
            }
        }
        token = __shfl_sync(mask, token, leader);
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar), "l"(token) : "memory");
        const double* cell = st.brick + (((i3 - m3) * 3 + (i2 - m2)) * 3 + (i1 - m1)) * 8;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double w = ((c & 2) ? d3 : e3) * ((c & 1) ? d2 : e2);
            const double w0 = w * e1, w1 = w * d1;
            const double* p = cell + ((c & 2) ? 72 : 0) + ((c & 1) ? 24 : 0);
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
                const double2 a = *reinterpret_cast<const double2*>(p + q), b = *reinterpret_cast<const double2*>(p + 8 + q);
                prims[q] = fma(w0, a.x, fma(w1, b.x, prims[q]));
                prims[q + 1] = fma(w0, a.y, fma(w1, b.y, prims[q + 1]));
            }
        }
        __syncwarp(mask);          // every lane has read its corners before the brick is overwritten again
    } else {
        const double* base = reinterpret_cast<const double*>(sn.cells) + (mb * sn.sb + i3 * sk + i2 * sj + (long)(i1 * 8));
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double w = ((c & 2) ? d3 : e3) * ((c & 1) ? d2 : e2);
            const double w0 = w * e1, w1 = w * d1;
            double a[8], b[8];
            load_cell_pair(base + ((c & 2) ? sk : 0) + ((c & 1) ? sj : 0), a, b);
#pragma unroll
            for (int q = 0; q < 8; q++) prims[q] = fma(w0, a[q], fma(w1, b[q], prims[q]));
        }
    }
}

__device__ __forceinline__ bool interp_prims_staged(const SnapshotView& sn, const double x[4], double prims[8], const CellStage& st)
{
    BlockGeom geo;
    int mb = locate_block_grid(sn, x, geo);
    if (mb < 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) prims[q] = 0.0;
        return false;
    }
    trilinear_staged(sn, mb, geo, x, prims, st);
    return true;
}
#endif

// Split form of the specialised f64 sampling path: locate (block, cell, weights) first, gather later -- lets a caller
// put independent work (and a prefetch of the four 128 B cell rows) between the address and its first use.
struct CellRef {
    const double* base;        // first of the eight corner cells
    double d1, d2, d3;         // fractional position inside the cell
};

MK_HD bool locate_cells_f64(const SnapshotView& sn, const double x[4], CellRef& ref)
{
    BlockGeom geo;
    int mb = locate_block_grid(sn, x, geo);
    if (mb < 0) return false;
    int i1, i2, i3;
    double inv[3];
    load_inv_dx(sn, mb, inv);
    cell_index(x[1], geo.v0[0], geo.dx[0], inv[0], true, i1, ref.d1);
    cell_index(x[2], geo.v0[1], geo.dx[1], inv[1], true, i2, ref.d2);
    cell_index(x[3], geo.v0[2], geo.dx[2], inv[2], true, i3, ref.d3);
    i1 = min(max(i1, 0), sn.ni); i2 = min(max(i2, 0), sn.nj); i3 = min(max(i3, 0), sn.nk);
    ref.base = reinterpret_cast<const double*>(sn.cells) + (mb * sn.sb + i3 * sn.sk + i2 * sn.sj + (long)(i1 * 8));
    return true;
}

MK_HD void prefetch_cells_f64(const SnapshotView& sn, const CellRef& ref)
{
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const double* p = ref.base + ((c & 2) ? sn.sk : 0) + ((c & 1) ? sn.sj : 0);
        asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
        asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 15));          // a cell pair may straddle two 128 B lines
    }
#else
    (void)sn; (void)ref;
#endif
}

MK_HD void gather_cells_f64(const SnapshotView& sn, const CellRef& ref, double prims[8])
{
    const double d1 = ref.d1, d2 = ref.d2, d3 = ref.d3;
    const double e1 = 1.0 - d1, e2 = 1.0 - d2, e3 = 1.0 - d3;
#pragma unroll
    for (int q = 0; q < 8; q++) prims[q] = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const double w = ((c & 2) ? d3 : e3) * ((c & 1) ? d2 : e2);
        const double w0 = w * e1, w1 = w * d1;
        double a[8], b[8];
        load_cell_pair(ref.base + ((c & 2) ? sn.sk : 0) + ((c & 1) ? sn.sj : 0), a, b);
#pragma unroll
        for (int q = 0; q < 8; q++) prims[q] = fma(w0, a[q], fma(w1, b[q], prims[q]));
    }
}

// Analytic torus primitives in canonical order; zero outside r <= r_out (the model's "domain").  Same formulas as
// mahakala_b200/synthetic.py::torus_fields (no waves); divisions, square roots and the two Gaussians use the
// MUFU-seeded FP64 helpers (<= few ulp; the exponentials flush results below 1e-307 to zero).
MK_HD bool torus_prims(const TorusParams& t, const double x[4], double prims[8])
{
    double R2 = fma(x[1], x[1], x[2] * x[2]);
    double R = fast_sqrt(R2) + 1e-12, r = fast_sqrt(fma(x[3], x[3], R2)) + 1e-12;
    if (!(r <= t.r_out)) {
#pragma unroll
        for (int q = 0; q < 8; q++) prims[q] = 0.0;
        return false;
    }
    double iR = fast_rcp(R);
    double q2 = t.R_in * iR;
    q2 = q2 * q2;
    double taper = fast_exp_neg((q2 * q2));                               // exp(-(R_in/R)^4)
    double zr = x[3] * iR;
    double gauss = fast_exp_neg(zr * zr * t.inv_2h2);                     // exp(-z^2 / (2 (h R)^2))
    double xr = R * t.inv_R0, plaw;
    if (t.p_is_three_halves) {
        double sx, isx;
        fast_sqrt_rsqrt(xr, sx, isx);
        plaw = isx * isx * isx;
    } else {
        plaw = pow(xr, -t.p);
    }
    double dens = t.dens_scale * plaw * gauss * taper;
    double eint = t.u0R0 * dens * fast_rcp(r);
    double s1R, vphi2;
    fast_sqrt_rsqrt(1. + R, s1R, vphi2);
    double vphi = 0.5 * vphi2;
    double bmag = fast_sqrt(eint * t.cB);
    double cx = x[1] * iR, cy = x[2] * iR;
    prims[0] = dens; prims[1] = eint;
    prims[2] = -vphi * cy; prims[3] = vphi * cx; prims[4] = 0.02 * x[3] * fast_rcp(1. + r);
    prims[5] = -bmag * cy; prims[6] = bmag * cx; prims[7] = 0.1 * bmag;
    return true;
}

// Snapshot kinds the fused render kernel is specialised for (its hot loop then carries only that kind's code: the
// kernel is sensitive to instruction-cache footprint and register pressure).  GENERIC handles everything at run time.
enum { SNAP_GENERIC = 0, SNAP_F64_GRID_POW2 = 1, SNAP_F32_GRID_POW2 = 2 };

MK_HD int snapshot_kind(const SnapshotView& sn)
{
    if (sn.source == 0 && sn.grid != nullptr && sn.dx_pow2) return sn.is_f32 ? SNAP_F32_GRID_POW2 : SNAP_F64_GRID_POW2;
    return SNAP_GENERIC;
}

template <int KIND>
MK_HD bool interp_prims_kind(const SnapshotView& sn, const double x[4], double prims[8]);

// prims in canonical order dens, eint, U1..3, B1..3; returns false (and zeros) outside the domain
MK_HD bool interp_prims(const SnapshotView& sn, const double x[4], double prims[8])
{
    if (sn.source == 1) return torus_prims(sn.torus, x, prims);
    BlockGeom geo;
    int mb = locate_block(sn, x, geo);
    if (mb < 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) prims[q] = 0.0;
        return false;
    }
    if (sn.is_f32) trilinear<float>(sn, mb, geo, x, prims);
    else trilinear<double>(sn, mb, geo, x, prims);
    return true;
}

template <int KIND>
MK_HD bool interp_prims_kind(const SnapshotView& sn, const double x[4], double prims[8])
{
    if (KIND == SNAP_GENERIC) return interp_prims(sn, x, prims);
    BlockGeom geo;
    int mb = locate_block_grid(sn, x, geo);
    if (mb < 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) prims[q] = 0.0;
        return false;
    }
    if (KIND == SNAP_F32_GRID_POW2) trilinear<float, 1>(sn, mb, geo, x, prims);
    else trilinear<double, 1>(sn, mb, geo, x, prims);
    return true;
}

struct FluidScalars {
    double dens, u, cos_pitch, kdotu, b;     // cos_pitch: clamped cosine; pitch_angle = acos(cos_pitch)
};

// athenak.py:760-792 with g = eta + f l l, g^-1 = eta - f l^m l^n in closed form (Kerr-Schild family).
MK_HD FluidScalars fluid_frame(double f, const double l[4], const double s[8],
                                                    const double prims[8], double cos_fallback)
{
    const double* U = prims + 2;
    const double* Bp = prims + 5;
    double alpha = sqrt(1.0 / (1.0 + f));                       // 1/sqrt(-g^tt), g^tt = -(1 + f)
    double lU = l[1] * U[0] + l[2] * U[1] + l[3] * U[2];
    double gamma = sqrt(1.0 + (U[0] * U[0] + U[1] * U[1] + U[2] * U[2]) + f * lU * lU);
    double ucon[4], ucov[4], bcon[4], bcov[4];
    ucon[0] = gamma / alpha;
    double ga = gamma * alpha;
#pragma unroll
    for (int i = 1; i < 4; i++) ucon[i] = U[i - 1] - ga * (f * l[i]);      // g^{0i} = f l_i
    double lu = l[0] * ucon[0] + l[1] * ucon[1] + l[2] * ucon[2] + l[3] * ucon[3];
    ucov[0] = -ucon[0] + f * l[0] * lu;
#pragma unroll
    for (int i = 1; i < 4; i++) ucov[i] = ucon[i] + f * l[i] * lu;
    bcon[0] = Bp[0] * ucov[1] + Bp[1] * ucov[2] + Bp[2] * ucov[3];
#pragma unroll
    for (int i = 1; i < 4; i++) bcon[i] = (Bp[i - 1] + ucon[i] * bcon[0]) / ucon[0];
    double lb = l[0] * bcon[0] + l[1] * bcon[1] + l[2] * bcon[2] + l[3] * bcon[3];
    bcov[0] = -bcon[0] + f * l[0] * lb;
#pragma unroll
    for (int i = 1; i < 4; i++) bcov[i] = bcon[i] + f * l[i] * lb;
    double kdotu = 0.0, kdotb = 0.0, bdotb = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        kdotu += s[4 + i] * ucov[i];
        kdotb += s[4 + i] * bcov[i];
        bdotb += bcon[i] * bcov[i];
    }
    double b = sqrt(bdotb);
    double c = kdotb / (fabs(kdotu) * b);
    if (c != c) c = cos_fallback;
    if (fabs(c) > 1.0) c = c / fabs(c);
    FluidScalars o;
    o.dens = prims[0]; o.u = prims[1]; o.cos_pitch = c; o.kdotu = kdotu; o.b = b;
    return o;
}

struct EmissionParams {
    double fluid_gamma, r_low, r_high, electron_gamma, ion_gamma;
    double Ne_unit, B_unit, L_unit;
    double sigma_cut;
    // physical constants are passed from the host module (mahakala_b200/constants.py) so that there is
    // exactly one table of digits
    double EE, CL, ME, MP, HPL;
    double two_11_12;           // 2^(11/12)
};

// transfer.py:56-86.  sin_pitch = sin(pitch_angle).  Returns (emissivity, absorptivity).
MK_HD void synchrotron(const EmissionParams& P, double Ne, double Theta_e, double B,
                                            double sin_pitch, double nu, int invariant, double rescale_nu,
                                            double& em_out, double& ab_out)
{
    const double PI = 3.141592653589793;
    double nuc = P.EE * B / (2. * PI * P.ME * P.CL);
    double th2 = Theta_e * Theta_e;
    double nus = (2. / 9.) * nuc * th2 * sin_pitch;
    double X = nu / nus;
    double x13 = cbrt(X);
    double var = exp(-x13);
    double term = sqrt(X) + P.two_11_12 * sqrt(x13);
    double em = Ne * nus * (term * term) / (2. * th2);
    em = em * var * 1.4142135623730951 * PI * (P.EE * P.EE) / (3.0 * P.CL);
    if (X > 1.e12) em = 0.0;
    if (Theta_e < 0.3) em = 0.0;
    double bx = P.HPL * nu / (P.ME * P.CL * P.CL * Theta_e);
    double den = (bx < 2.e-3) ? bx / 24. * (24. + bx * (12. + bx * (4. + bx))) : exp(bx) - 1.0;
    double B_nu = (2. * P.HPL * (nu * nu * nu) / den) / (P.CL * P.CL);
    double ab = em / B_nu;
    if (invariant) {
        double rn = nu * rescale_nu;
        em = em / (rn * rn);
        ab = ab * rn;
    }
    em_out = (em != em) ? 0.0 : em;
    ab_out = (ab != ab) ? 0.0 : ab;
}

// images.py:87-102 + electrons.py:46-50: everything between the fluid scalars and (Ne, Theta_e, B)
MK_HD void plasma_state(const EmissionParams& P, const FluidScalars& fs, double& Ne,
                                             double& Theta_e, double& Bg, double& sigma)
{
    double bsq = fs.b * fs.b;
    double beta = fs.u * (P.fluid_gamma - 1.) / bsq / 0.5;
    sigma = bsq / fs.dens;
    double b2 = beta * beta;
    double T_ratio = (P.r_high * b2 + P.r_low) / (1. + b2);
    double t_e = (P.CL * P.CL) * (P.MP * fs.u * (P.electron_gamma - 1.) * (P.ion_gamma - 1.));
    t_e /= fs.dens * ((P.ion_gamma - 1.) + (P.electron_gamma - 1.) * T_ratio);
    Theta_e = t_e / (P.ME * P.CL * P.CL);
    Ne = P.Ne_unit * fs.dens;
    Bg = P.B_unit * fs.b;
}

// ---------------------------------------------------------------------------------------------------------
// Fused fast path (render kernel): fluid frame -> plasma state -> invariant j_nu, alpha_nu for NF frequencies.
//
// The reference evaluates the whole chain in IEEE arithmetic and maps every NaN to zero at the end
// (transfer.py:83-84).  Working through its special cases: the emissivity is non-zero only when
//     dens > 0, u > 0, b.b > 0, k.u < 0, |cos(pitch)| < 1, sigma <= cut, Theta_e >= 0.3 and 0 < X <= 1e12
// (anything else yields 0, NaN -> 0, or the explicit zeroing of transfer.py:71-72 / images.py:116-118, and
// the absorptivity em / B_nu is zero or NaN -> 0 with it).  Inside that region every operand is a positive
// normal number, so divisions become MUFU-seeded reciprocals and sqrt(X) = (X^(1/6))^3; results agree with
// the IEEE chain to ~1e-15 relative.  Returns false when the sample contributes nothing.
// ---------------------------------------------------------------------------------------------------------
struct EmissionConsts {      // derived once per launch on the host from EmissionParams
    double g1x2;             // 2 (fluid_gamma - 1)
    double theta_fac;        // (MP / ME) (electron_gamma - 1) (ion_gamma - 1)
    double igm1, egm1;       // ion_gamma - 1, electron_gamma - 1
    double k_nus;            // (2/9) EE / (2 pi ME CL) * B_unit
    double k_em;             // sqrt(2) pi EE^2 / (6 CL) * Ne_unit
    double k_bx;             // HPL / (ME CL^2)
    double k_ab;             // CL^2 / (2 HPL)
    // frequency ratios relative to nu_obs[0]: X, bx and 1/nu of frequency f follow from those of frequency 0
    // by one multiplication, so cbrt / sqrt / reciprocal are evaluated once per sample, not once per frequency
    double nu0, inv_nu0;     // nu_obs[0] and its reciprocal
    double ratio[8];         // nu_f / nu_0
    double iratio3[8];       // (nu_0 / nu_f)^3
    double c13[8];           // cbrt(nu_f / nu_0)
    double c16[8];           // (nu_f / nu_0)^(1/6)
};

__host__ __device__ inline EmissionConsts make_emission_consts(const EmissionParams& P, const double* nu_obs = nullptr,
                                                              int nfreq = 0)
{
    const double PI = 3.141592653589793;
    EmissionConsts c;
    c.g1x2 = 2.0 * (P.fluid_gamma - 1.0);
    c.theta_fac = (P.MP / P.ME) * (P.electron_gamma - 1.0) * (P.ion_gamma - 1.0);
    c.igm1 = P.ion_gamma - 1.0;
    c.egm1 = P.electron_gamma - 1.0;
    c.k_nus = (2. / 9.) * P.EE / (2. * PI * P.ME * P.CL) * P.B_unit;
    c.k_em = 1.4142135623730951 * PI * (P.EE * P.EE) / (6.0 * P.CL) * P.Ne_unit;
    c.k_bx = P.HPL / (P.ME * P.CL * P.CL);
    c.k_ab = (P.CL * P.CL) / (2.0 * P.HPL);
    c.nu0 = (nu_obs && nfreq > 0) ? nu_obs[0] : 1.0;
    c.inv_nu0 = 1.0 / c.nu0;
    for (int f = 0; f < 8; f++) {
        double r = (nu_obs && f < nfreq) ? nu_obs[f] / c.nu0 : 1.0;
        c.ratio[f] = r;
        c.iratio3[f] = 1.0 / (r * r * r);
        c.c13[f] = cbrt(r);
        c.c16[f] = sqrt(c.c13[f]);
    }
    return c;
}

// Frame scalars of one sample: what the emission chain needs from the fluid-frame algebra of athenak.py:760-786.
struct FrameScalars {
    double kdotu, kdotb, bsq;
};

// ... for the Kerr-Schild family g = eta + f l l, g^-1 = eta - f l^m l^n in closed form
MK_HD FrameScalars frame_kerr_schild(double f, const double l[4], const double s[8], const double prims[8])
{
    const double* U = prims + 2;
    const double* Bp = prims + 5;
    double sq1f, alpha;                                          // sqrt(1+f), 1/sqrt(1+f) = lapse
    quick_sqrt_rsqrt(1.0 + f, sq1f, alpha);
    double lU = fma(l[1], U[0], fma(l[2], U[1], l[3] * U[2]));
    double gamma = quick_sqrt(fma(f * lU, lU, 1.0 + fma(U[0], U[0], fma(U[1], U[1], U[2] * U[2]))));
    double ucon[4], ucov[4], bcon[4], bcov[4];
    ucon[0] = gamma * sq1f;
    double gaf = gamma * alpha * f;
#pragma unroll
    for (int i = 1; i < 4; i++) ucon[i] = fma(-gaf, l[i], U[i - 1]);
    double flu = f * (ucon[0] + fma(l[1], ucon[1], fma(l[2], ucon[2], l[3] * ucon[3])));
    ucov[0] = flu - ucon[0];
#pragma unroll
    for (int i = 1; i < 4; i++) ucov[i] = fma(flu, l[i], ucon[i]);
    bcon[0] = fma(Bp[0], ucov[1], fma(Bp[1], ucov[2], Bp[2] * ucov[3]));
    double iu0 = fast_rcp(ucon[0]);
#pragma unroll
    for (int i = 1; i < 4; i++) bcon[i] = fma(ucon[i], bcon[0], Bp[i - 1]) * iu0;
    double flb = f * (bcon[0] + fma(l[1], bcon[1], fma(l[2], bcon[2], l[3] * bcon[3])));
    bcov[0] = flb - bcon[0];
#pragma unroll
    for (int i = 1; i < 4; i++) bcov[i] = fma(flb, l[i], bcon[i]);
    FrameScalars o;
    o.kdotu = fma(s[4], ucov[0], fma(s[5], ucov[1], fma(s[6], ucov[2], s[7] * ucov[3])));
    o.kdotb = fma(s[4], bcov[0], fma(s[5], bcov[1], fma(s[6], bcov[2], s[7] * bcov[3])));
    o.bsq = fma(bcon[0], bcov[0], fma(bcon[1], bcov[1], fma(bcon[2], bcov[2], bcon[3] * bcov[3])));
    return o;
}

// ... for an arbitrary spacetime given its covariant and contravariant metric at the sample (user-registered
// plugins): athenak.py:760-786 term by term
MK_HD FrameScalars frame_generic(const double g[4][4], const double gi[4][4], const double s[8], const double prims[8])
{
    const double* U = prims + 2;
    const double* Bp = prims + 5;
    double alpha = 1.0 / sqrt(-gi[0][0]);
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) q = fma(g[i + 1][j + 1] * U[i], U[j], q);
    double gamma = sqrt(1.0 + q);
    double ucon[4], ucov[4], bcon[4], bcov[4];
    ucon[0] = gamma / alpha;
#pragma unroll
    for (int i = 1; i < 4; i++) ucon[i] = fma(-gamma * alpha, gi[0][i], U[i - 1]);
#pragma unroll
    for (int i = 0; i < 4; i++) ucov[i] = fma(g[i][0], ucon[0], fma(g[i][1], ucon[1], fma(g[i][2], ucon[2], g[i][3] * ucon[3])));
    bcon[0] = fma(Bp[0], ucov[1], fma(Bp[1], ucov[2], Bp[2] * ucov[3]));
#pragma unroll
    for (int i = 1; i < 4; i++) bcon[i] = fma(ucon[i], bcon[0], Bp[i - 1]) / ucon[0];
#pragma unroll
    for (int i = 0; i < 4; i++) bcov[i] = fma(g[i][0], bcon[0], fma(g[i][1], bcon[1], fma(g[i][2], bcon[2], g[i][3] * bcon[3])));
    FrameScalars o;
    o.kdotu = fma(s[4], ucov[0], fma(s[5], ucov[1], fma(s[6], ucov[2], s[7] * ucov[3])));
    o.kdotb = fma(s[4], bcov[0], fma(s[5], bcov[1], fma(s[6], bcov[2], s[7] * bcov[3])));
    o.bsq = fma(bcon[0], bcov[0], fma(bcon[1], bcov[1], fma(bcon[2], bcov[2], bcon[3] * bcov[3])));
    return o;
}

// `sink(fq, em, ab)` receives the invariant emissivity and absorptivity of frequency fq as soon as they are known, so
// that a multi-frequency caller can fold them into its accumulators without holding 2 NF values in registers.
template <int NF, class Sink>
MK_HD bool emission_from_frame(const EmissionParams& P, const EmissionConsts& C, const FrameScalars& fr,
                               const double prims[8], Sink&& sink)
{
    // Straight-line code: every validity test only feeds the final select (invalid lanes may carry NaN/inf
    // through the arithmetic, which is harmless on the GPU).  Early exits would make the compiler duplicate
    // the copies of all loop-carried registers on every exit edge (~200 MOVs per sample in SASS).
    const double dens = prims[0], u = prims[1];
    // Theta_e = (const > 0) u / dens must reach 0.3, so dens and u have the same sign; both positive in any physical
    // snapshot, both negative (Ne < 0: negative j and alpha in the reference too) kept for input-for-input parity
    bool valid = ((dens > 0.0) & (u > 0.0)) | ((dens < 0.0) & (u < 0.0));
    const double kdotu = fr.kdotu, kdotb = fr.kdotb, bsq = fr.bsq;
    valid &= (bsq > 0.0) & (kdotu < 0.0);
    double b, ib;
    quick_sqrt_rsqrt(bsq, b, ib);
    // cos(pitch), athenak.py:789-791.  The reference clamps it to [-1, 1]; a clamped value gives sin = 0, hence
    // nu_s = 0, X = inf and zero emissivity, which is what "sin2 > 0 fails" yields here without the clamp
    // (NaN also fails the test)
    double rn = -kdotu;
    double irn = fast_rcp(rn);
    double c = kdotb * ib * irn;
    double sin2 = (1.0 - c) * (1.0 + c);
    valid &= (sin2 > 0.0);
    double sinp = quick_sqrt(sin2);
    // ---- plasma state (images.py:87-102, electrons.py:46-50) ----
    double idens = fast_rcp(dens);
    double sigma = bsq * idens;
    valid &= !(sigma > P.sigma_cut);
    double beta = C.g1x2 * u * (ib * ib);
    double b2 = beta * beta;
    double T_ratio = fma(P.r_high, b2, P.r_low) * fast_rcp(1.0 + b2);
    double Theta = C.theta_fac * u * idens * fast_rcp(fma(C.egm1, T_ratio, C.igm1));
    valid &= (Theta >= 0.3);
    // ---- synchrotron (transfer.py:56-81), invariant form ----
    double th2 = Theta * Theta;
    double nus = C.k_nus * b * th2 * sinp;
    double inus = fast_rcp(nus);
    double ith = fast_rcp(Theta);
    double pref = C.k_em * dens * nus * (ith * ith);
    // quantities of frequency 0; frequency f scales them by powers of nu_f / nu_0
    double nu0 = rn * C.nu0;
    double X0 = nu0 * inus;
    double ix13_0;
    double x13_0 = fast_cbrt_pos(X0, ix13_0);
    if (valid && !((X0 > 1e-30) & (X0 < 1e30))) x13_0 = cbrt(X0);     // outside the float-seeded range (rare)
    double x16_0 = quick_sqrt(x13_0);
    double bx0 = C.k_bx * nu0 * ith;
    // invariant rescaling (transfer.py:77-80): nu * rescale_nu = (-k.u nu_f) / nu_f = -k.u for every frequency
    double irn2 = irn * irn;
    double inu0 = irn * C.inv_nu0;                               // 1 / (-k.u nu_0)
    double kab0 = C.k_ab * (inu0 * inu0 * inu0);
#pragma unroll
    for (int fq = 0; fq < NF; fq++) {
        double X = X0 * C.ratio[fq];
        double x13 = x13_0 * C.c13[fq];
        double x16 = x16_0 * C.c16[fq];
        double term = fma(x16 * x16, x16, P.two_11_12 * x16);
        double e = pref * (term * term) * fast_exp_neg(x13);
        double bx = bx0 * C.ratio[fq];
        double den = (bx < 2.e-3) ? bx * (1. / 24.) * fma(bx, fma(bx, 4. + bx, 12.), 24.) : exp(bx) - 1.0;
        // (k_ab / nu_0^3) (nu_0 / nu_f)^3 first: with frequencies decades apart the product e den k_ab / nu_0^3
        // would underflow before the ratio brings it back
        double a = (e * den) * (kab0 * C.iratio3[fq]);
        e = e * irn2;
        a = a * rn;
        bool ok = valid & (X <= 1.e12) & (e == e) & (a == a);
        sink(fq, ok ? e : 0.0, ok ? a : 0.0);
    }
    return valid;
}

// Kerr-Schild front end (the fused kernel of the built-in spacetime and the emission probe)
template <int NF, class Sink>
MK_HD bool emission_fast(const EmissionParams& P, const EmissionConsts& C, double f, const double l[4],
                         const double s[8], const double prims[8], const double* nu_obs, const double* inv_nu_obs,
                         Sink&& sink)
{
    (void)nu_obs; (void)inv_nu_obs;
    return emission_from_frame<NF>(P, C, frame_kerr_schild(f, l, s, prims), prims, sink);
}

}  // namespace mk
