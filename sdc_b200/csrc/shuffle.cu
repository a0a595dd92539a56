// shuffle.cu -- hash partitioning for the multi-GPU shuffle (K11 in SURVEY.md).
//
// Replaces the partition step of the reference's (dead) MPI shuffle -- per-row destination rank and
// per-destination counts, sdc/shuffle_utils.py:119-163 -- the exchange itself (c_alltoall of the counts,
// c_alltoallv per column, sdc/distributed_api.py:439-466, a memcpy stub in
// sdc/transport/hpat_transport_single_process.cpp:106-119) is NCCL all-to-all in sdc_b200/distributed.py.
//
//   dest[i]  = mix64(canonical key) % nparts   (same canonicalisation as groupby / join: -0.0 == +0.0,
//              one NaN; NaN keys still get a destination -- the local kernels decide to skip them)
//   counts   = rows per destination (int64, unlike the reference's int32 counts, distributed_api.py:122)
//   perm     = stable grouping of the row numbers by destination: one 8-bit onesweep pass, so every
//              destination's rows stay in input order (left-order joins and first-appearance groupby
//              survive the shuffle)
// Columns are then packed with sdcb200_gather(perm) into contiguous per-destination send buffers.
#include "radix_sort.cuh"

namespace sdcb200 {
namespace {

constexpr int kPartThreads = 256;
constexpr int kMaxParts = 256;

__host__ __device__ inline unsigned long long pmix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <bool KEY_F64>
__global__ void __launch_bounds__(kPartThreads) partition_dest_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                                                      unsigned nparts, uint8_t* __restrict__ dest,
                                                                      unsigned long long* __restrict__ counts) {
    __shared__ unsigned hist[kMaxParts];
    for (int i = threadIdx.x; i < kMaxParts; i += kPartThreads) hist[i] = 0;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)kPartThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kPartThreads) {
        unsigned long long b = keys[i];
        if (KEY_F64) {
            const unsigned long long mag = b & 0x7fffffffffffffffull;
            if (mag > 0x7ff0000000000000ull) b = 0x7ff8000000000000ull;
            else if (mag == 0) b = 0;
        }
        // the TOP bits choose the rank, so the low bits the local hash tables index with stay uniform
        const unsigned d = (unsigned)(((pmix64(b) >> 32) * (unsigned long long)nparts) >> 32);
        dest[i] = (uint8_t)d;
        atomicAdd(&hist[d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)nparts; i += kPartThreads)
        if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" int32_t sdcb200_hash_partition(int32_t key_dtype, const void* keys, int64_t n, int32_t nparts, int64_t* perm_out,
                                          int64_t* counts_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(n >= 0 && nparts >= 1 && nparts <= kMaxParts, "n / nparts (<= 256)");
    SDC_ARG_CHECK(counts_out != nullptr && (n == 0 || (keys && perm_out)), "null pointer");
    SDC_CUDA_TRY(cudaMemsetAsync(counts_out, 0, (size_t)nparts * 8, s));
    if (n == 0) return SDCB200_OK;
    Scratch dest;
    SDC_TRY(dest.alloc((size_t)n, s));
    int64_t grid = ceil_div(n, kPartThreads * 8);
    if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
    if (grid < 1) grid = 1;
    const unsigned long long* k = reinterpret_cast<const unsigned long long*>(keys);
    unsigned long long* c = reinterpret_cast<unsigned long long*>(counts_out);
    if (key_dtype == SDCB200_F64)
        partition_dest_kernel<true><<<(unsigned)grid, kPartThreads, 0, s>>>(k, n, (unsigned)nparts, dest.as<uint8_t>(), c);
    else
        partition_dest_kernel<false><<<(unsigned)grid, kPartThreads, 0, s>>>(k, n, (unsigned)nparts, dest.as<uint8_t>(), c);
    SDC_CHECK_LAUNCH();
    SDC_TRY(argsort_device(SDCB200_U8, dest.p, n, true, perm_out, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));      // `dest` goes out of scope
    return SDCB200_OK;
}

// out[i] = cnt[i] ? sum[i] / cnt[i] : NaN -- the combine step of a distributed mean (partial sums and
// counts are exchanged, the quotient is taken once: nanmean, sdc/functions/numpy_like.py:775-792)
namespace sdcb200 { namespace {
__global__ void combine_mean_kernel(const double* __restrict__ sum, const long long* __restrict__ cnt, int64_t n,
                                    double* __restrict__ out) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = cnt[i] ? sum[i] / (double)cnt[i] : nan;
}
}}

extern "C" int32_t sdcb200_combine_mean(const double* sum, const int64_t* cnt, int64_t n, double* out, void* stream) {
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    int64_t grid = ceil_div(n, 256);
    if (grid > (int64_t)sm_count() * 4) grid = (int64_t)sm_count() * 4;
    combine_mean_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(sum, reinterpret_cast<const long long*>(cnt), n, out);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}
