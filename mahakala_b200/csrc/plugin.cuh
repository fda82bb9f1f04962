// Host-side hooks into the run-time metric registry (plugin.cu).
#pragma once
#include "common.cuh"
#include "integrate_kernel.cuh"

namespace mk {
int plugin_integrate(int metric_id, double bhspin, IntegrateArgs& A, cudaStream_t stream);
int plugin_elementwise(int metric_id, double bhspin, const char* kernel, void** extra_args, int n_extra, long n,
                       cudaStream_t stream);
}
