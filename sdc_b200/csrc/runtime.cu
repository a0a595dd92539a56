// runtime.cu -- error state, launch counter, allocation and copy helpers of libsdc_b200.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace sdcb200 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

static int32_t ensure_pool() {
    static std::once_flag once[64];
    int dev = 0;
    SDC_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return SDCB200_OK;
    cudaError_t err = cudaSuccess;
    std::call_once(once[dev], [&] {
        cudaMemPool_t pool;
        err = cudaDeviceGetDefaultMemPool(&pool, dev);
        if (err == cudaSuccess) {
            uint64_t thr = UINT64_MAX;  // keep freed blocks cached in the pool
            err = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    });
    SDC_CUDA_TRY(err);
    return SDCB200_OK;
}

int32_t dev_alloc(void** p, size_t bytes, cudaStream_t s) {
    SDC_TRY(ensure_pool());
    SDC_CUDA_TRY(cudaMallocAsync(p, bytes, s));
    return SDCB200_OK;
}

int32_t dev_free(void* p, cudaStream_t s) {
    if (!p) return SDCB200_OK;
    SDC_CUDA_TRY(cudaFreeAsync(p, s));
    return SDCB200_OK;
}

}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

const char* sdcb200_last_error(void) { return g_err; }

const char* sdcb200_version(void) { return "sdc_b200 0.1 (sm_100a)"; }

int64_t sdcb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int32_t sdcb200_device_info(int32_t* sms, int64_t* hbm_bytes, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    SDC_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SDC_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (sms) *sms = prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return SDCB200_OK;
}

int32_t sdcb200_host_alloc(void** p, int64_t bytes) {
    SDC_ARG_CHECK(p != nullptr && bytes >= 0, "host_alloc");
    SDC_CUDA_TRY(cudaHostAlloc(p, (size_t)(bytes ? bytes : 16), cudaHostAllocDefault));
    return SDCB200_OK;
}

int32_t sdcb200_host_free(void* p) {
    if (p) SDC_CUDA_TRY(cudaFreeHost(p));
    return SDCB200_OK;
}

int32_t sdcb200_dev_alloc(void** p, int64_t bytes, void* stream) {
    SDC_ARG_CHECK(p != nullptr && bytes >= 0, "dev_alloc");
    return dev_alloc(p, (size_t)(bytes ? bytes : 16), (cudaStream_t)stream);
}

int32_t sdcb200_dev_free(void* p, void* stream) { return dev_free(p, (cudaStream_t)stream); }

int32_t sdcb200_memcpy_h2d(void* dst, const void* src, int64_t bytes, void* stream) {
    SDC_ARG_CHECK(bytes >= 0, "memcpy_h2d");
    if (bytes) SDC_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return SDCB200_OK;
}

int32_t sdcb200_memcpy_d2h(void* dst, const void* src, int64_t bytes, void* stream) {
    SDC_ARG_CHECK(bytes >= 0, "memcpy_d2h");
    if (bytes) SDC_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return SDCB200_OK;
}

int32_t sdcb200_stream_sync(void* stream) {
    SDC_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return SDCB200_OK;
}

}  // extern "C"
