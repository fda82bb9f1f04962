// Library glue of libmahakala_b200.so: error string, device info, persistent-kernel queue counters,
// the DFMA peak microbenchmark and small utility kernels (radius_cal, rhs on a bundle).
#include <cstdarg>
#include <cstring>
#include <mutex>
#include "common.cuh"
#include "integrate.cuh"
#include "ks_metric.cuh"
#include "metric_plugin.cuh"
#include "plugin.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// A small pool of device counters per device; each persistent launch takes one slot and zeroes it on
// its own stream, so back-to-back launches on one stream are ordered and never share a live counter.
unsigned int* queue_counter(cudaStream_t stream, int slot)
{
    static unsigned int* pool[64] = {nullptr};
    static unsigned next[64] = {0};
    static std::mutex mtx;
    constexpr int SLOTS = 64;
    std::lock_guard<std::mutex> lock(mtx);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("cudaGetDevice failed"); return nullptr; }
    if (!pool[dev]) {
        if (cudaMalloc(&pool[dev], SLOTS * 64 * sizeof(unsigned int)) != cudaSuccess) {
            set_error("cudaMalloc of the queue counters failed");
            return nullptr;
        }
    }
    (void)slot;
    unsigned int* p = pool[dev] + (size_t)(next[dev]++ % SLOTS) * 64;     // 256 B apart
    if (cudaMemsetAsync(p, 0, 64 * sizeof(unsigned int), stream) != cudaSuccess) {
        set_error("cudaMemsetAsync of the queue counter failed");
        return nullptr;
    }
    return p;
}

// ---- DFMA peak: 8 independent FMA chains per thread, full occupancy --------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;       // never true; keeps the chains alive
}

__global__ void radius_kernel(KerrSchild g, const double* x, long n, long stride, double* r)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* p = x + i * stride;
    // literal radius_cal with IEEE sqrt (utility path, not the integrator's hot loop)
    double R = sqrt(p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    double w = R * R - g.aa;
    r[i] = sqrt((w + sqrt(w * w + 4.0 * g.aa * (p[3] * p[3]))) / 2.0);
}

template <class Metric>
__global__ void rhs_kernel(Metric g, const double* state, long n, double* out)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[8], acc[4];
#pragma unroll
    for (int m = 0; m < 8; m++) s[m] = state[i * 8 + m];
    g.accel(s, s + 4, acc);
#pragma unroll
    for (int m = 0; m < 4; m++) { out[i * 8 + m] = s[4 + m]; out[i * 8 + 4 + m] = acc[m]; }
}

__global__ void fast_math_probe_kernel(const double* x, long n, double* rcp, double* sq, double* rsq)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, r;
    fast_sqrt_rsqrt(x[i], s, r);
    rcp[i] = fast_rcp(x[i]);
    sq[i] = s;
    rsq[i] = r;
}

__global__ void transcendental_probe_kernel(const double* x, long n, double* expneg, double* cbrt_out, double* icbrt)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    expneg[i] = fast_exp_neg(x[i]);
    cbrt_out[i] = fast_cbrt_pos(x[i], r);
    icbrt[i] = r;
}

template <class Metric>
__global__ void rk4_kernel(Metric g, const double* state, const double* dt, long n, double* out)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[8], o[8];
#pragma unroll
    for (int m = 0; m < 8; m++) s[m] = state[i * 8 + m];
    rk4_step(g, s, dt[i], o);
#pragma unroll
    for (int m = 0; m < 8; m++) out[i * 8 + m] = o[m];
}

template <class Metric>
__global__ void metric_kernel(Metric g, const double* x, long n, double* gout, double* giout)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[4], gc[4][4], gi[4][4];
#pragma unroll
    for (int m = 0; m < 4; m++) p[m] = x[i * 4 + m];
    g.metric_cov_con(p, gc, gi);
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (gout) gout[i * 16 + a * 4 + b] = gc[a][b];
            if (giout) giout[i * 16 + a * 4 + b] = gi[a][b];
        }
}

}  // namespace mk
using namespace mk;

extern "C" int mk_abi_version(void) { return MK_ABI_VERSION; }

extern "C" const char* mk_last_error_string(void) { return g_error; }

extern "C" int mk_device_info(int* sms, int* cc_major, int* cc_minor, int* sm_clock_khz, long* total_mem)
{
    int dev = 0;
    MK_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    MK_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (sm_clock_khz) { int k = 0; cudaDeviceGetAttribute(&k, cudaDevAttrClockRate, dev); *sm_clock_khz = k; }
    if (total_mem) *total_mem = (long)p.totalGlobalMem;
    return 0;
}

extern "C" int mk_measure_fp64_peak(int iters, double* tflops_out, double* ms_out)
{
    MK_REQUIRE(iters > 0, "iters must be positive");
    double* d = nullptr;
    MK_CUDA_CHECK(cudaMalloc(&d, 64));
    int blocks = sm_count() * 8;
    cudaEvent_t e0, e1;
    MK_CUDA_CHECK(cudaEventCreate(&e0));
    MK_CUDA_CHECK(cudaEventCreate(&e1));
    dfma_peak_kernel<<<blocks, 256>>>(d, iters / 8 + 1, 0.999999, 1e-9);     // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        MK_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    MK_CUDA_CHECK(cudaGetLastError());
    double fmas = (double)blocks * 256.0 * (double)iters * 16.0 * 8.0;
    if (tflops_out) *tflops_out = 2.0 * fmas / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return 0;
}

extern "C" int mk_radius_cal(double bhspin, const double* x, long n, long stride, double* r, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(x && r, "null pointer");
    KerrSchild g; g.set_spin(bhspin);
    radius_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, x, n, stride, r);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_rhs(int metric_id, double bhspin, const double* state, long n, double* out, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(state && out, "null pointer");
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (metric_id == MK_METRIC_KERR_SCHILD) {
        KerrSchild g; g.set_spin(bhspin);
        rhs_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(g, state, n, out);
    } else if (metric_id == MK_METRIC_KERR_SCHILD_DUAL) {
        DualMetric<KerrSchildFn> g; g.fn.a = bhspin; g.rH = 0;
        rhs_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(g, state, n, out);
    } else if (metric_id >= MK_METRIC_PLUGIN_BASE) {
        void* extra[] = {&state, &n, &out};
        return plugin_elementwise(metric_id, bhspin, "mk_plugin_rhs", extra, 3, n, (cudaStream_t)stream);
    } else {
        set_error("unknown metric id %d", metric_id);
        return 2;
    }
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

#define MK_DISPATCH_METRIC(metric_id, bhspin, CALL)                                                   \
    if (metric_id == MK_METRIC_KERR_SCHILD) {                                                          \
        KerrSchild g; g.set_spin(bhspin);  \
        CALL;                                                                                          \
    } else if (metric_id == MK_METRIC_KERR_SCHILD_DUAL) {                                              \
        DualMetric<KerrSchildFn> g; g.fn.a = bhspin; g.rH = 1.0 + sqrt(1.0 - bhspin * bhspin);         \
        CALL;                                                                                          \
    } else {                                                                                           \
        set_error("unknown metric id %d", metric_id);                                                  \
        return 2;                                                                                      \
    }

extern "C" int mk_rk4_step(int metric_id, double bhspin, const double* state, const double* dt, long n,
                           double* out, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(state && dt && out, "null pointer");
    if (metric_id >= MK_METRIC_PLUGIN_BASE) {
        void* extra[] = {&state, &dt, &n, &out};
        return plugin_elementwise(metric_id, bhspin, "mk_plugin_rk4", extra, 4, n, (cudaStream_t)stream);
    }
    unsigned blocks = (unsigned)((n + 127) / 128);
    MK_DISPATCH_METRIC(metric_id, bhspin, (rk4_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(g, state, dt, n, out)));
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_metric(int metric_id, double bhspin, const double* x, long n, double* gc, double* gi, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(x != nullptr, "null pointer");
    if (metric_id >= MK_METRIC_PLUGIN_BASE) {
        void* extra[] = {&x, &n, &gc, &gi};
        return plugin_elementwise(metric_id, bhspin, "mk_plugin_metric", extra, 4, n, (cudaStream_t)stream);
    }
    unsigned blocks = (unsigned)((n + 127) / 128);
    MK_DISPATCH_METRIC(metric_id, bhspin, (metric_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(g, x, n, gc, gi)));
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- peer memory (CUDA IPC) for the shared tile queue / in-kernel gather -----------------------------
extern "C" int mk_ipc_alloc(long bytes, void** ptr, unsigned char* handle64)
{
    MK_REQUIRE(bytes > 0 && ptr && handle64, "bad arguments");
    void* p = nullptr;
    MK_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
    MK_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return 1;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    *ptr = p;
    return 0;
}

extern "C" int mk_ipc_open(const unsigned char* handle64, void** ptr)
{
    MK_REQUIRE(handle64 && ptr, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    MK_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int mk_ipc_close(void* ptr)
{
    if (ptr) MK_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int mk_ipc_free(void* ptr)
{
    if (ptr) MK_CUDA_CHECK(cudaFree(ptr));
    return 0;
}

extern "C" int mk_fast_math_probe(const double* x, long n, double* rcp, double* sq, double* rsq, void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(x && rcp && sq && rsq, "null pointer");
    fast_math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, rcp, sq, rsq);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int mk_transcendental_probe(const double* x, long n, double* expneg, double* cbrt_out, double* icbrt,
                                       void* stream)
{
    if (n <= 0) return 0;
    MK_REQUIRE(x && expneg && cbrt_out && icbrt, "null pointer");
    transcendental_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, expneg, cbrt_out, icbrt);
    MK_CUDA_CHECK(cudaGetLastError());
    return 0;
}
