// Run-time registered spacetimes: NVRTC compiles the user's metric functor together with the integrate kernel
// body (plugin_tu.cuh) for sm_100a; the cubin is loaded with cudaLibraryLoadData and launched like the
// built-in kernels.  The reference's equivalent is swapping the module-level metric()/imetric()
// (/root/reference/mahakala/geodesics.py:88-104, :304-305, :339-347) and letting jax.jacfwd differentiate it.
#include <nvrtc.h>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "common.cuh"
#include "integrate_kernel.cuh"
#include "plugin.cuh"
#include "../../include/mahakala_b200.h"

namespace mk {

struct PluginBlobHost { double params[8]; };

struct Plugin {
    std::string name;
    std::vector<char> cubin;
    double params[8];
    cudaLibrary_t lib[64];
    bool loaded[64];
};

static std::vector<Plugin*> g_plugins;
static std::mutex g_mutex;

static Plugin* find_plugin(int metric_id)
{
    int k = metric_id - MK_METRIC_PLUGIN_BASE;
    if (k < 0 || k >= (int)g_plugins.size()) {
        set_error("unknown metric id %d (%d run-time metrics registered)", metric_id, (int)g_plugins.size());
        return nullptr;
    }
    return g_plugins[k];
}

static int get_kernel(Plugin* p, const char* name, cudaKernel_t* k)
{
    int dev = 0;
    MK_CUDA_CHECK(cudaGetDevice(&dev));
    MK_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    if (!p->loaded[dev]) {
        MK_CUDA_CHECK(cudaLibraryLoadData(&p->lib[dev], p->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
        p->loaded[dev] = true;
    }
    MK_CUDA_CHECK(cudaLibraryGetKernel(k, p->lib[dev], name));
    return 0;
}

static void fill_blob(const Plugin* p, double bhspin, PluginBlobHost& b)
{
    memcpy(b.params, p->params, sizeof b.params);
    b.params[0] = bhspin;
}

int plugin_integrate(int metric_id, double bhspin, IntegrateArgs& A, cudaStream_t stream)
{
    Plugin* p = find_plugin(metric_id);
    if (!p) return 2;
    cudaKernel_t k;
    const char* name = A.pages ? "mk_plugin_integrate_paged" : (A.S ? "mk_plugin_integrate_padded" : "mk_plugin_integrate_final");
    if (int rc = get_kernel(p, name, &k)) return rc;
    PluginBlobHost b;
    fill_blob(p, bhspin, b);
    long blocks = (long)sm_count() * 4;
    long need = ((A.npx + 31) / 32 + 3) / 4;
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    if (A.chunk_div > 0 || A.ray_order) {
        set_error("run-time registered spacetimes integrate on one GPU's own queue (no shared queue / ray order)");
        return 2;
    }
    void* args[] = {&b, &A};
    MK_CUDA_CHECK(cudaLaunchKernel((const void*)k, dim3((unsigned)blocks), dim3(128), args, 0, stream));
    return 0;
}

int plugin_render(int metric_id, double bhspin, const void* render_args, size_t args_bytes, long npatches,
                  cudaStream_t stream)
{
    (void)args_bytes;
    Plugin* p = find_plugin(metric_id);
    if (!p) return 2;
    cudaKernel_t k;
    if (int rc = get_kernel(p, "mk_plugin_render", &k)) return rc;
    PluginBlobHost b;
    fill_blob(p, bhspin, b);
    long blocks = (long)sm_count() * 2;                 // __launch_bounds__(128, 2) in plugin_tu.cuh
    long need = (npatches + 3) / 4;
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    void* args[] = {&b, const_cast<void*>(render_args)};
    MK_CUDA_CHECK(cudaLaunchKernel((const void*)k, dim3((unsigned)blocks), dim3(128), args, 0, stream));
    return 0;
}

int plugin_elementwise(int metric_id, double bhspin, const char* kernel, void** extra_args, int n_extra, long n,
                       cudaStream_t stream)
{
    Plugin* p = find_plugin(metric_id);
    if (!p) return 2;
    cudaKernel_t k;
    if (int rc = get_kernel(p, kernel, &k)) return rc;
    PluginBlobHost b;
    fill_blob(p, bhspin, b);
    void* args[8];
    args[0] = &b;
    for (int i = 0; i < n_extra; i++) args[1 + i] = extra_args[i];
    MK_CUDA_CHECK(cudaLaunchKernel((const void*)k, dim3((unsigned)((n + 127) / 128)), dim3(128), args, 0, stream));
    return 0;
}

}  // namespace mk
using namespace mk;

extern "C" int mk_register_metric(const char* name, const char* source, const char* include_dir, int* metric_id,
                                  char* log, long log_capacity)
{
    MK_REQUIRE(name && source && include_dir && metric_id, "null pointer");
    std::string src = "#include \"metric_plugin.cuh\"\nusing namespace mk;\n#line 1 \"user_metric.cu\"\n";
    src += source;
    src += "\n#include \"plugin_tu.cuh\"\n";
    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, src.c_str(), "mk_user_metric.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) {
        set_error("nvrtcCreateProgram failed: %s", nvrtcGetErrorString(r));
        return 1;
    }
    std::string inc = std::string("-I") + include_dir;
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", inc.c_str(), "-lineinfo", "-default-device"};
    r = nvrtcCompileProgram(prog, 5, opts);
    size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    std::string plog(log_size, '\0');
    if (log_size > 1) nvrtcGetProgramLog(prog, &plog[0]);
    if (log && log_capacity > 0) {
        strncpy(log, plog.c_str(), (size_t)log_capacity - 1);
        log[log_capacity - 1] = '\0';
    }
    if (r != NVRTC_SUCCESS) {
        set_error("compilation of metric '%s' failed: %s\n%.700s", name, nvrtcGetErrorString(r), plog.c_str());
        nvrtcDestroyProgram(&prog);
        return 3;
    }
    size_t n = 0;
    nvrtcGetCUBINSize(prog, &n);
    Plugin* p = new Plugin();
    p->name = name;
    p->cubin.resize(n);
    nvrtcGetCUBIN(prog, p->cubin.data());
    nvrtcDestroyProgram(&prog);
    memset(p->params, 0, sizeof p->params);
    memset(p->loaded, 0, sizeof p->loaded);
    std::lock_guard<std::mutex> lock(g_mutex);
    g_plugins.push_back(p);
    *metric_id = MK_METRIC_PLUGIN_BASE + (int)g_plugins.size() - 1;
    return 0;
}

extern "C" int mk_metric_set_params(int metric_id, const double* params8)
{
    Plugin* p = find_plugin(metric_id);
    if (!p) return 2;
    MK_REQUIRE(params8 != nullptr, "null pointer");
    memcpy(p->params, params8, sizeof p->params);
    return 0;
}

extern "C" int mk_initial_condition_metric(int metric_id, double bhspin, const double* s0_x, const double* s0_v,
                                           long n, double* s0, void* stream)
{
    if (metric_id < MK_METRIC_PLUGIN_BASE) return mk_initial_condition(bhspin, s0_x, s0_v, n, s0, stream);
    if (n <= 0) return 0;
    MK_REQUIRE(s0 && s0_x && s0_v, "null pointer");
    void* extra[] = {&s0_x, &s0_v, &n, &s0};
    return plugin_elementwise(metric_id, bhspin, "mk_plugin_nullify", extra, 4, n, (cudaStream_t)stream);
}
