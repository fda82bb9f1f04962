// Host-side hooks into the run-time metric registry (plugin.cu).
#pragma once
#include "common.cuh"
#include "integrate_kernel.cuh"

namespace mk {
int plugin_integrate(int metric_id, double bhspin, IntegrateArgs& A, cudaStream_t stream);
// fused render with a registered spacetime: args = a RenderArgs (render_kernel.cuh), passed opaquely so that this
// header does not pull the kernel body into every translation unit
int plugin_render(int metric_id, double bhspin, const void* render_args, size_t args_bytes, long npatches,
                  cudaStream_t stream);
int plugin_elementwise(int metric_id, double bhspin, const char* kernel, void** extra_args, int n_extra, long n,
                       cudaStream_t stream);
}
