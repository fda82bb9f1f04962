// Warp-specialised long-patch render kernels (render_pipeline.cuh): __global__ instantiations for 1..8 observing
// frequencies x snapshot kind x {128-thread shared-SM CTAs, 512-thread SM-exclusive CTAs} and their launcher.  Own
// translation unit so that it compiles next to render.cu.
#include "common.cuh"
#include "render_pipeline.cuh"
#include "snapshot.cuh"

namespace mk {

template <int NF, int KIND, int GROUPS>
__global__ void __launch_bounds__(PIPE_THREADS * GROUPS, GROUPS == 1 ? 4 : 1) render_pipeline_kernel(const KerrSchild g, const RenderArgs A)
{
    render_pipeline_body<NF, KIND, GROUPS>(g, A);
}

template <int NF, int KIND, int GROUPS>
static int launch_pipeline_groups(const KerrSchild& g, const RenderArgs& A, long npatches, int max_ctas, cudaStream_t stream)
{
    const size_t smem = GROUPS * (size_t)PipeLayout<NF>::GROUP_BYTES;
    // per device and cheap: set on every launch rather than cached per process
    MK_CUDA_CHECK(cudaFuncSetAttribute(render_pipeline_kernel<NF, KIND, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_pipeline_kernel<NF, KIND, GROUPS>, PIPE_THREADS * GROUPS, smem));
    if (per_sm < 1) per_sm = 1;
    long blocks = (long)sm_count() * per_sm;        // one patch per group at a time
    const int working = (GROUPS == 1) ? 1 : A.pipe_groups;
    const long need = (npatches + working - 1) / working;
    if (need < blocks) blocks = need;
    if (max_ctas > 0 && max_ctas < blocks) blocks = max_ctas;
    if (blocks < 1) blocks = 1;
    render_pipeline_kernel<NF, KIND, GROUPS><<<(unsigned)blocks, PIPE_THREADS * GROUPS, smem, stream>>>(g, A);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int NF, int KIND>
static int launch_pipeline_kind(const KerrSchild& g, const RenderArgs& A, long npatches, int exclusive, int max_ctas, cudaStream_t stream)
{
    if (exclusive) {
        RenderArgs B = A;
        B.pipe_groups = exclusive > PIPE_GROUPS_EXCLUSIVE ? PIPE_GROUPS_EXCLUSIVE : exclusive;
        return launch_pipeline_groups<NF, KIND, PIPE_GROUPS_EXCLUSIVE>(g, B, npatches, max_ctas, stream);
    }
    return launch_pipeline_groups<NF, KIND, 1>(g, A, npatches, max_ctas, stream);
}

template <int NF>
static int launch_pipeline_nf(const KerrSchild& g, const RenderArgs& A, long npatches, int exclusive, int max_ctas, cudaStream_t stream)
{
    switch (snapshot_kind(A.sn)) {
        case SNAP_F64_GRID_POW2: return launch_pipeline_kind<NF, SNAP_F64_GRID_POW2>(g, A, npatches, exclusive, max_ctas, stream);
        case SNAP_F32_GRID_POW2: return launch_pipeline_kind<NF, SNAP_F32_GRID_POW2>(g, A, npatches, exclusive, max_ctas, stream);
        default: return launch_pipeline_kind<NF, SNAP_GENERIC>(g, A, npatches, exclusive, max_ctas, stream);
    }
}

int launch_render_pipeline(const KerrSchild& g, const RenderArgs& A, int nfreq, long npatches, int exclusive, int max_ctas,
                           cudaStream_t stream)
{
    switch (nfreq) {
        case 1: return launch_pipeline_nf<1>(g, A, npatches, exclusive, max_ctas, stream);
        case 2: return launch_pipeline_nf<2>(g, A, npatches, exclusive, max_ctas, stream);
        case 3: return launch_pipeline_nf<3>(g, A, npatches, exclusive, max_ctas, stream);
        case 4: return launch_pipeline_nf<4>(g, A, npatches, exclusive, max_ctas, stream);
        case 5: return launch_pipeline_nf<5>(g, A, npatches, exclusive, max_ctas, stream);
        case 6: return launch_pipeline_nf<6>(g, A, npatches, exclusive, max_ctas, stream);
        case 7: return launch_pipeline_nf<7>(g, A, npatches, exclusive, max_ctas, stream);
        default: return launch_pipeline_nf<8>(g, A, npatches, exclusive, max_ctas, stream);
    }
}

}  // namespace mk
