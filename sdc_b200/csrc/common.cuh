// common.cuh -- shared host/device helpers for libsdc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/sdc_b200.h"

namespace sdcb200 {

// ---------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

#define SDC_CUDA_TRY(expr)                                                         \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            ::sdcb200::set_error("%s failed: %s (%s:%d)", #expr,                  \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);     \
            return (_e == cudaErrorMemoryAllocation) ? SDCB200_ERR_NOMEM          \
                                                     : SDCB200_ERR_CUDA;          \
        }                                                                          \
    } while (0)

#define SDC_TRY(expr)                                                              \
    do {                                                                           \
        int32_t _s = (expr);                                                       \
        if (_s != SDCB200_OK) return _s;                                           \
    } while (0)

#define SDC_CHECK_LAUNCH()                                                         \
    do {                                                                           \
        ::sdcb200::count_launch();                                                 \
        SDC_CUDA_TRY(cudaGetLastError());                                          \
    } while (0)

#define SDC_ARG_CHECK(cond, msg)                                                   \
    do {                                                                           \
        if (!(cond)) {                                                             \
            ::sdcb200::set_error("bad argument: %s (%s)", msg, #cond);            \
            return SDCB200_ERR_ARG;                                                \
        }                                                                          \
    } while (0)

// stream-ordered scratch allocation (default mempool, release threshold = max
// so freed blocks stay cached between calls).
int32_t dev_alloc(void** p, size_t bytes, cudaStream_t s);
int32_t dev_free(void* p, cudaStream_t s);

// RAII scratch buffer for use inside a C-ABI function body.
struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    Scratch() = default;
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    int32_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        return dev_alloc(&p, bytes ? bytes : 16, stream);
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
    ~Scratch() { if (p) dev_free(p, s); }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// hash_mix64 picks hash-table slots and partition buckets from its TOP bits (partitioned join / groupby): one xor-shift and
// one multiply by an odd constant (Fibonacci hashing over the folded key) -- 7 instructions instead of the 19 of a full
// murmur finaliser, which the partition pass pays per row.  The shuffle's destination rank comes from a murmur finaliser
// over the same key (shuffle.cu): a different function, so the keys a rank owns after a hash shuffle still spread over
// all buckets / slots.
__host__ __device__ inline unsigned long long mix64_fin(unsigned long long x) {      // murmur3 finaliser
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__host__ __device__ inline unsigned long long hash_mix64(unsigned long long x) {
    x ^= x >> 32;
    return x * 0x9e3779b97f4a7c15ull;
}

// -------------------------------------------------------------- device side
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted on `bar`.  16-B aligned src/dst, bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global bulk store (bulk async-group completion).
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- streaming 128-bit global access (no L1 allocation) ----
__device__ __forceinline__ int4 ld_stream_v4(const void* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v2f64(double* p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// ---- L2 eviction-priority hints (createpolicy + ld/st ... L2::cache_hint) ----
// evict_first for data streamed once (probe keys, emitted pairs), evict_last for the random-access
// table that should stay L2-resident while the streams pass through.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ld_hint_u32(const uint32_t* a, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ unsigned long long ld_hint_u64(const unsigned long long* a, uint64_t pol) {
    unsigned long long v;
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ ulonglong2 ld_hint_v2u64(const void* a, uint64_t pol) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
                 : "=l"(v.x), "=l"(v.y) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ int4 ld_hint_v4(const void* p, uint64_t pol) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_hint_v2s64(void* a, long long x, long long y, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1,%2}, %3;" ::"l"(a), "l"(x), "l"(y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint_s64(void* a, long long x, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(a), "l"(x), "l"(pol) : "memory");
}

__device__ __forceinline__ double shfl_up_f64(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

__device__ __forceinline__ bool is_finite_f64(double v) {
    // exponent field != all ones
    return ((__double2hiint(v) & 0x7ff00000) != 0x7ff00000);
}

#endif  // __CUDACC__

}  // namespace sdcb200
