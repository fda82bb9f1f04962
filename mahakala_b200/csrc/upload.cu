// Host -> device upload of large PAGEABLE arrays (the interior cell arrays an AthenaK loader returns) through two
// pinned staging buffers: several host threads copy chunk k+1 into one buffer while the DMA engine moves chunk k out
// of the other.  A plain cudaMemcpy from pageable memory is staged by the driver on a single thread (5-8 GB/s on the
// bench hosts); pinning a multi-GB array first costs more than the copy.  Snapshot ingestion is a "next" row of the
// hot path (the reference re-uploads the ghost-padded array on every call, athenak.py:693).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"
#include "../../include/mahakala_b200.h"

namespace {
constexpr size_t STAGE_BYTES = 32ull << 20;
struct Stage {
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool ok = false;
};
Stage g_stage;
std::mutex g_mtx;

bool stage_init()
{
    if (g_stage.ok) return true;
    for (int i = 0; i < 2; i++) {
        if (cudaHostAlloc(&g_stage.buf[i], STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&g_stage.done[i], cudaEventDisableTiming) != cudaSuccess) return false;
    }
    g_stage.ok = true;
    return true;
}

void parallel_copy(char* dst, const char* src, size_t n, int threads)
{
    if (threads <= 1 || n < (4u << 20)) { std::memcpy(dst, src, n); return; }
    std::vector<std::thread> pool;
    size_t per = ((n + threads - 1) / threads + 4095) & ~size_t(4095);
    for (int t = 0; t < threads; t++) {
        size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
        if (hi > lo) pool.emplace_back([=] { std::memcpy(dst + lo, src + lo, hi - lo); });
    }
    for (auto& th : pool) th.join();
}
}  // namespace

// Strategy "register": pin the caller's pages in place (cudaHostRegister), let the DMA engine read them directly,
// unpin.  No host-side copy at all -- on hosts whose memcpy bandwidth is a few GB/s (virtualised bench boxes) the
// staging copy is the bottleneck of the staged strategy.
static int upload_registered(char* dst, const char* src, size_t total, cudaStream_t stream)
{
    const size_t page = 4096;
    const char* lo = (const char*)((uintptr_t)src & ~(uintptr_t)(page - 1));
    const char* hi = (const char*)(((uintptr_t)(src + total) + page - 1) & ~(uintptr_t)(page - 1));
    cudaError_t e = cudaHostRegister((void*)lo, (size_t)(hi - lo), cudaHostRegisterReadOnly);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaHostRegister((void*)lo, (size_t)(hi - lo), cudaHostRegisterDefault);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }          // caller falls back to staging
    cudaError_t c = cudaMemcpyAsync(dst, src, total, cudaMemcpyHostToDevice, stream);
    if (c == cudaSuccess) c = cudaStreamSynchronize(stream);
    cudaHostUnregister((void*)lo);
    if (c != cudaSuccess) { mk::set_error("registered upload failed: %s", cudaGetErrorString(c)); return 1; }
    return 0;
}

extern "C" int mk_upload_pageable(void* dst_device, const void* src_host, long bytes, int host_threads, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MK_REQUIRE(bytes >= 0, "negative size");
    if (bytes == 0) return 0;
    MK_REQUIRE(dst_device && src_host, "null pointer");
    // host_threads < 0 selects the in-place registration strategy (falls back to staging if the pages cannot be pinned)
    if (host_threads < 0) {
        int rc = upload_registered((char*)dst_device, (const char*)src_host, (size_t)bytes, stream);
        if (rc >= 0) return rc;
        host_threads = 0;
    }
    std::lock_guard<std::mutex> lock(g_mtx);
    if (!stage_init()) { mk::set_error("pinned staging buffers could not be allocated"); return 1; }
    int threads = host_threads > 0 ? host_threads : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    const char* src = (const char*)src_host;
    char* dst = (char*)dst_device;
    size_t off = 0, total = (size_t)bytes;
    for (int k = 0; off < total; k++) {
        int b = k & 1;
        size_t len = std::min(STAGE_BYTES, total - off);
        if (k >= 2) MK_CUDA_CHECK(cudaEventSynchronize(g_stage.done[b]));      // the DMA out of this buffer has finished
        parallel_copy((char*)g_stage.buf[b], src + off, len, threads);
        MK_CUDA_CHECK(cudaMemcpyAsync(dst + off, g_stage.buf[b], len, cudaMemcpyHostToDevice, stream));
        MK_CUDA_CHECK(cudaEventRecord(g_stage.done[b], stream));
        off += len;
    }
    // the staging buffers are reused by the next call: wait for the last two DMAs
    MK_CUDA_CHECK(cudaEventSynchronize(g_stage.done[0]));
    MK_CUDA_CHECK(cudaEventSynchronize(g_stage.done[1]));
    return 0;
}
