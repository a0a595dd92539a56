// scan.cuh -- device-wide exclusive scan (reduce-then-scan over tiles) shared by join.cu / strings.cu.
// Everything here is a template or static, so each translation unit gets its own copy.
#pragma once
#include "common.cuh"

namespace sdcb200 {
namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;       // 2048 entries per CTA

// ---- device-wide exclusive scan, reduce-then-scan over tiles of kScanTile entries ----
// LOAD(i) -> u64 value of entry i; STORE(i, exclusive prefix, value)
struct ArrayLoad {
    const unsigned long long* p;
    __device__ unsigned long long operator()(int64_t i) const { return p[i]; }
};
struct ArrayStore {
    unsigned long long* p;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long) const { p[i] = excl; }
};

__device__ __forceinline__ unsigned long long block_excl_scan_u64(unsigned long long v, unsigned long long* total,
                                                                  unsigned long long* wsum /* [8] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        const unsigned long long x = wsum[w];
        if (w < warp) base += x;
        tot += x;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

template <typename LOAD>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(LOAD load, int64_t n, unsigned long long* tsum) {
    __shared__ unsigned long long wsum[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    unsigned long long v = 0;
#pragma unroll
    for (int e = 0; e < kScanItems; ++e) {
        const int64_t i = base + e * kScanThreads + threadIdx.x;
        if (i < n) v += load(i);
    }
    unsigned long long tot;
    block_excl_scan_u64(v, &tot, wsum);
    if (threadIdx.x == 0) tsum[blockIdx.x] = tot;
}

// one CTA of 1024 threads: exclusive scan of the tile sums in place; total -> tsum[nt].  The CTA walks the array in
// chunks of 4096 entries, thread t holding entries 4t .. 4t+3 of the chunk (coalesced), with a running carry -- the
// earlier thread-contiguous split read the array with a stride of nt / 1024 entries between neighbouring threads and
// took 100 us for 48 829 entries.
constexpr int kScanTilesThreads = 1024;
static __global__ void __launch_bounds__(kScanTilesThreads) scan_tiles_kernel(unsigned long long* tsum, int64_t nt) {
    __shared__ unsigned long long wsum[kScanTilesThreads / 32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    for (int64_t c0 = 0; c0 < nt; c0 += 4 * kScanTilesThreads) {
        const int64_t i0 = c0 + 4 * (int64_t)threadIdx.x;
        unsigned long long v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (i0 + e < nt) ? tsum[i0 + e] : 0ull;
        const unsigned long long s = v[0] + v[1] + v[2] + v[3];
        unsigned long long incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long base = s_carry, tot = 0;
#pragma unroll
        for (int w = 0; w < kScanTilesThreads / 32; ++w) {
            const unsigned long long x = wsum[w];
            if (w < warp) base += x;
            tot += x;
        }
        unsigned long long run = base + incl - s;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (i0 + e < nt) tsum[i0 + e] = run;
            run += v[e];
        }
        __syncthreads();                       // every thread has read wsum and s_carry
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tsum[nt] = s_carry;
}

template <typename LOAD, typename STORE>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(LOAD load, STORE store, int64_t n,
                                                                  const unsigned long long* __restrict__ tsum) {
    __shared__ unsigned long long wsum[kScanThreads / 32];
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;   // thread-contiguous items
    unsigned long long v[kScanItems], s = 0;
#pragma unroll
    for (int e = 0; e < kScanItems; ++e) {
        v[e] = (i0 + e < n) ? load(i0 + e) : 0ull;
        s += v[e];
    }
    unsigned long long tot;
    unsigned long long ex = tsum[blockIdx.x] + block_excl_scan_u64(s, &tot, wsum);
#pragma unroll
    for (int e = 0; e < kScanItems; ++e) {
        if (i0 + e < n) store(i0 + e, ex, v[e]);
        ex += v[e];
    }
}

template <typename LOAD, typename STORE>
int32_t device_exclusive_scan(LOAD load, STORE store, int64_t n, unsigned long long* tsum /* [nt + 1] */, cudaStream_t s) {
    const int64_t nt = ceil_div(n, kScanTile);
    scan_reduce_kernel<<<(unsigned)nt, kScanThreads, 0, s>>>(load, n, tsum);
    SDC_CHECK_LAUNCH();
    scan_tiles_kernel<<<1, kScanTilesThreads, 0, s>>>(tsum, nt);
    SDC_CHECK_LAUNCH();
    scan_apply_kernel<<<(unsigned)nt, kScanThreads, 0, s>>>(load, store, n, tsum);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}


}  // namespace
}  // namespace sdcb200
