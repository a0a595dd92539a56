// sort.cu -- onesweep LSD radix sort / stable argsort / gather (K1, K2, K3 in SURVEY.md).
//
// Replaces sdc.concurrent_sort (sdc/native/sort.cpp:42-155 tbb::parallel_sort,
// sdc/native/stable_sort.cpp:91-344 TBB merge sort) and numpy_like.take
// (sdc/functions/numpy_like.py:1359-1388).  Order semantics = the reference comparators
// (sdc/native/utils.cpp:160-182): NaN last for ascending AND descending, -0.0 == +0.0, and --
// because an LSD radix sort is stable per digit -- ties keep their input order in both
// directions, which is what parallel_stable_argsort_u64<t> guarantees.
//
// Structure (one read of the keys for all digit histograms, then one read + one write per
// 8-bit digit that is not constant over the column):
//   hist_kernel      per-CTA shared-memory histograms (native u32 ATOMS) -> global
//   scan_kernel      exclusive scan per digit position + "digit is constant" flags
//   onesweep_kernel  per tile: warp-ballot multisplit ranks, decoupled look-back across tiles
//                    for the per-digit global offsets, shared-memory staged scatter so each
//                    digit's run is written as one contiguous (coalesced) span
// The digit is taken from a canonical key computed on the fly from the raw bits, so the rows
// that move are the original values (a stable keys-only sort keeps -0.0 / +0.0 / NaN payloads
// in input order exactly as a comparison sort with these comparators would).
#include "radix_sort.cuh"

namespace sdcb200 {

int dtype_size(int dtype) {
    switch (dtype) {
        case SDCB200_I8: case SDCB200_U8: return 1;
        case SDCB200_I16: case SDCB200_U16: return 2;
        case SDCB200_I32: case SDCB200_U32: case SDCB200_F32: return 4;
        case SDCB200_I64: case SDCB200_U64: case SDCB200_F64: return 8;
    }
    return 0;
}

namespace {

enum { KIND_UINT = 0, KIND_SINT = 1, KIND_FLOAT = 2 };

int dtype_kind(int dtype) {
    switch (dtype) {
        case SDCB200_I8: case SDCB200_I16: case SDCB200_I32: case SDCB200_I64: return KIND_SINT;
        case SDCB200_F32: case SDCB200_F64: return KIND_FLOAT;
    }
    return KIND_UINT;
}

template <typename U> struct FloatMask { static constexpr U inf = 0; };
template <> struct FloatMask<uint32_t> { static constexpr uint32_t inf = 0x7f800000u; };
template <> struct FloatMask<uint64_t> { static constexpr uint64_t inf = 0x7ff0000000000000ull; };

// raw bits -> unsigned key whose ascending order is the requested order
template <typename U>
__device__ __forceinline__ U canon_key(U b, int kind, bool desc) {
    constexpr U sign = U(U(1) << (8 * sizeof(U) - 1));
    U k;
    if (sizeof(U) >= 4 && kind == KIND_FLOAT) {
        const U mag = U(b & U(~sign));
        if (mag > FloatMask<U>::inf) return U(~U(0));   // NaN: last in both directions
        if (mag == 0) b = 0;                            // -0.0 == +0.0
        k = (b & sign) ? U(~b) : U(b | sign);
    } else if (kind == KIND_SINT) {
        k = U(b ^ sign);
    } else {
        k = b;
    }
    return desc ? U(~k) : k;
}

constexpr int kHistThreads = 512;

template <typename U>
__global__ void __launch_bounds__(kHistThreads) hist_kernel(const U* __restrict__ keys, int64_t n, int kind, int desc,
                                                            unsigned long long* __restrict__ ghist) {
    constexpr int NP = sizeof(U);
    __shared__ unsigned sh[NP][256];
    for (int i = threadIdx.x; i < NP * 256; i += kHistThreads) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kHistThreads;
    for (int64_t i = (int64_t)blockIdx.x * kHistThreads + threadIdx.x; i < n; i += stride) {
        const U k = canon_key<U>(keys[i], kind, desc != 0);
#pragma unroll
        for (int p = 0; p < NP; ++p) atomicAdd(&sh[p][(unsigned)(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NP * 256; i += kHistThreads) {
        const unsigned v = (&sh[0][0])[i];
        if (v) atomicAdd(&ghist[i], (unsigned long long)v);
    }
}

// one CTA of 256 threads per digit position: exclusive scan of the 256 bins; flag constant digits
__global__ void __launch_bounds__(256) scan_kernel(const unsigned long long* __restrict__ ghist,
                                                   unsigned long long* __restrict__ gbase, int* __restrict__ trivial,
                                                   unsigned long long n) {
    __shared__ unsigned long long warp_tot[8];
    const int p = blockIdx.x, d = threadIdx.x, lane = d & 31, w = d >> 5;
    const unsigned long long v = ghist[p * 256 + d];
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    unsigned long long add = 0;
    for (int i = 0; i < w; ++i) add += warp_tot[i];
    gbase[p * 256 + d] = add + x - v;
    const int is_all = __syncthreads_or(v == n);
    if (d == 0) trivial[p] = is_all;
}

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;

// rows per thread: tiles of 4096 rows keep the number of tiles in flight (and so the length of the
// look-back chains) down without costing a resident CTA
template <typename U> struct Items { static constexpr int v = 16; };

struct NoPay {};

// look-back status word: bits 0-1 flag (1 = tile count, 2 = inclusive prefix), bits 2-9 epoch (the pass
// number + 1: entries of earlier passes are simply stale, so the buffer is zeroed once per sort, not once
// per pass), bits 10.. value
__device__ __forceinline__ unsigned long long lb_pack(unsigned long long v, unsigned flag, unsigned epoch) {
    return (v << 10) | ((unsigned long long)epoch << 2) | flag;
}

// 16-byte rows stage 80 KB per tile: two CTAs per SM fit, so the register budget is that of two
template <typename U, typename P, bool HAS_PAY, int KIND, bool DESC, int ITEMS>
__global__ void __launch_bounds__(kSortThreads, (HAS_PAY && sizeof(U) + sizeof(P) > 12) ? 2 : 3)
onesweep_kernel(const U* __restrict__ kin, U* __restrict__ kout, const P* __restrict__ pin, P* __restrict__ pout,
                int64_t n, int shift, int iota,
                const unsigned long long* __restrict__ gbase, unsigned long long* lookback, unsigned epoch) {
    constexpr int TILE = kSortThreads * ITEMS;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned* wh = reinterpret_cast<unsigned*>(smem);                        // [kSortWarps][256]
    unsigned* tile_off = wh + kSortWarps * 256;                              // [256]
    long long* digit_dst = reinterpret_cast<long long*>(tile_off + 256);    // [256]
    unsigned* warp_tot = reinterpret_cast<unsigned*>(digit_dst + 256);      // [8]
    U* skeys = reinterpret_cast<U*>(warp_tot + 16);
    P* spay = reinterpret_cast<P*>(skeys + TILE);
    unsigned char* sdig = reinterpret_cast<unsigned char*>(spay + (HAS_PAY ? TILE : 0));   // [TILE] digit of the staged row

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // tile = blockIdx.x: CTAs are dispatched in blockIdx order, so a tile's predecessors are resident or done
    const unsigned tile = blockIdx.x;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int tile_n = (int)((n - tile_base) < (int64_t)TILE ? (n - tile_base) : (int64_t)TILE);

    // ---- load (warp-striped: item i of lane l sits at warp_base + i*32 + l) ----
    U key[ITEMS];
    unsigned short rank[ITEMS];
    const int wbase = w * 32 * ITEMS;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int pos = wbase + i * 32 + lane;
        key[i] = pos < tile_n ? kin[tile_base + pos] : U(0);
    }
    // the payload rides along in registers: its loads are in flight while the tile is ranked
    P pay[HAS_PAY ? ITEMS : 1];
    if (HAS_PAY) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int pos = wbase + i * 32 + lane;
            pay[i] = iota ? P(tile_base + pos) : (pos < tile_n ? pin[tile_base + pos] : P(0));
        }
    }
    for (int i = tid; i < kSortWarps * 256; i += kSortThreads) wh[i] = 0;
    __syncthreads();
    // ---- rank inside the warp: ballot multisplit, stable in lane order ----
    unsigned* mywh = wh + w * 256;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int pos = wbase + i * 32 + lane;
        const bool valid = pos < tile_n;
        const unsigned d = (unsigned)(canon_key<U>(key[i], KIND, DESC) >> shift) & 255u;
        unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? m : ~m;
        }
        const unsigned prior = mywh[d];
        __syncwarp();
        if (valid && (peers & lt) == 0) mywh[d] = prior + __popc(peers);   // lowest peer lane updates
        __syncwarp();
        rank[i] = (unsigned short)(prior + __popc(peers & lt));
    }
    __syncthreads();
    // ---- per digit (thread d): prefix over warps, tile count published at once, scan over digits ----
    volatile unsigned long long* lb = lookback;
    unsigned long long cnt;
    {
        const int d = tid;
        unsigned run = 0;
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ++ww) {
            const unsigned t = wh[ww * 256 + d];
            wh[ww * 256 + d] = run;
            run += t;
        }
        cnt = run;
        lb[(size_t)tile * 256 + d] = lb_pack(cnt, tile == 0 ? 2u : 1u, epoch);
        unsigned x = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) warp_tot[w] = x;
        __syncthreads();
        unsigned add = 0;
        for (int i = 0; i < w; ++i) add += warp_tot[i];
        tile_off[d] = add + x - run;
    }
    __syncthreads();
    // ---- stage the tile in digit order (needs only tile-local offsets; predecessors publish meanwhile) ----
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int pos = wbase + i * 32 + lane;
        if (pos < tile_n) {
            const unsigned d = (unsigned)(canon_key<U>(key[i], KIND, DESC) >> shift) & 255u;
            const unsigned p = tile_off[d] + mywh[d] + rank[i];
            skeys[p] = key[i];
            sdig[p] = (unsigned char)d;
            if (HAS_PAY) spay[p] = pay[i];
        }
    }
    // ---- decoupled look-back per digit ----
    {
        const int d = tid;
        unsigned long long excl = 0;
        if (tile != 0) {
            long long t = (long long)tile - 1;
            while (true) {
                unsigned long long v;
                do { v = lb[(size_t)t * 256 + d]; } while ((v & 3ull) == 0 || (unsigned)((v >> 2) & 255ull) != epoch);
                excl += v >> 10;
                if ((v & 3ull) == 2ull) break;
                --t;
            }
            lb[(size_t)tile * 256 + d] = lb_pack(excl + cnt, 2u, epoch);
        }
        digit_dst[d] = (long long)(gbase[d] + excl) - (long long)tile_off[d];
    }
    __syncthreads();
    // ---- scatter: consecutive staged slots of one digit go to consecutive global slots ----
    for (int p = tid; p < tile_n; p += kSortThreads) {
        const long long dst = digit_dst[sdig[p]] + p;
        kout[dst] = skeys[p];
        if (HAS_PAY) pout[dst] = spay[p];
    }
}

template <typename P>
__global__ void iota_kernel(P* out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (P)i;
}

__global__ void widen_u32_kernel(const uint32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int64_t)in[i];
}

template <typename T>
__global__ void gather_kernel(const T* __restrict__ src, const int64_t* __restrict__ idx, T* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}

template <typename U, typename P, bool HAS_PAY, int ITEMS>
int32_t sort_impl_items(const U* src, U* bufA, U* bufB, const P* psrc, P* pA, P* pB, int64_t n, int kind, bool iota,
                  bool desc, const void** rkeys, const void** rpay, cudaStream_t s) {
    constexpr int NP = sizeof(U);
    constexpr int TILE = kSortThreads * ITEMS;
    const int64_t ntiles = ceil_div(n, TILE);
    Scratch meta, lb;
    // meta: ghist[NP*256] u64 | gbase[NP*256] u64 | counters[NP] u32 | trivial[NP] i32
    const size_t meta_bytes = (size_t)NP * 256 * 8 * 2 + NP * 4 + NP * 4;
    SDC_TRY(meta.alloc(meta_bytes, s));
    SDC_TRY(lb.alloc((size_t)ntiles * 256 * 8, s));
    unsigned long long* ghist = meta.as<unsigned long long>();
    unsigned long long* gbase = ghist + NP * 256;
    unsigned* counters = reinterpret_cast<unsigned*>(gbase + NP * 256);
    int* trivial = reinterpret_cast<int*>(counters + NP);
    SDC_CUDA_TRY(cudaMemsetAsync(meta.p, 0, meta_bytes, s));
    SDC_CUDA_TRY(cudaMemsetAsync(lb.p, 0, (size_t)ntiles * 256 * 8, s));
    int hgrid = (int)(ceil_div(n, kHistThreads * 8) < (int64_t)sm_count() * 4 ? ceil_div(n, kHistThreads * 8) : (int64_t)sm_count() * 4);
    if (hgrid < 1) hgrid = 1;
    hist_kernel<U><<<hgrid, kHistThreads, 0, s>>>(src, n, kind, desc ? 1 : 0, ghist);
    SDC_CHECK_LAUNCH();
    scan_kernel<<<NP, 256, 0, s>>>(ghist, gbase, trivial, (unsigned long long)n);
    SDC_CHECK_LAUNCH();
    int htriv[8];
    SDC_CUDA_TRY(cudaMemcpyAsync(htriv, trivial, NP * sizeof(int), cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));

    using KernT = void (*)(const U*, U*, const P*, P*, int64_t, int, int, const unsigned long long*, unsigned long long*, unsigned);
    KernT kern = nullptr;
    if (kind == KIND_FLOAT) kern = desc ? (KernT)onesweep_kernel<U, P, HAS_PAY, KIND_FLOAT, true, ITEMS> : (KernT)onesweep_kernel<U, P, HAS_PAY, KIND_FLOAT, false, ITEMS>;
    else if (kind == KIND_SINT) kern = desc ? (KernT)onesweep_kernel<U, P, HAS_PAY, KIND_SINT, true, ITEMS> : (KernT)onesweep_kernel<U, P, HAS_PAY, KIND_SINT, false, ITEMS>;
    else kern = desc ? (KernT)onesweep_kernel<U, P, HAS_PAY, KIND_UINT, true, ITEMS> : (KernT)onesweep_kernel<U, P, HAS_PAY, KIND_UINT, false, ITEMS>;
    const size_t smem = (size_t)kSortWarps * 256 * 4 + 256 * 4 + 256 * 8 + 16 * 4 + (size_t)TILE * sizeof(U) +
                        (HAS_PAY ? (size_t)TILE * sizeof(P) : 0) + (size_t)TILE;
    SDC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const U* kcur = src;
    const P* pcur = psrc;
    bool first = true;
    int flip = 0;
    for (int p = 0; p < NP; ++p) {
        if (htriv[p]) continue;
        U* kdst = flip ? bufB : bufA;
        P* pdst = flip ? pB : pA;
        kern<<<(unsigned)ntiles, kSortThreads, smem, s>>>(kcur, kdst, pcur, pdst, n, 8 * p,
                                                          (first && iota) ? 1 : 0, gbase + p * 256,
                                                          lb.as<unsigned long long>(), (unsigned)(p + 1));
        SDC_CHECK_LAUNCH();
        kcur = kdst;
        pcur = pdst;
        first = false;
        flip ^= 1;
    }
    if (first && HAS_PAY && iota) {   // every digit constant: the identity permutation
        iota_kernel<P><<<sm_count() * 4, 256, 0, s>>>(pA, n);
        SDC_CHECK_LAUNCH();
        pcur = pA;
    }
    *rkeys = kcur;
    *rpay = pcur;
    return SDCB200_OK;
}

template <typename U, typename P, bool HAS_PAY>
int32_t sort_impl(const U* src, U* bufA, U* bufB, const P* psrc, P* pA, P* pB, int64_t n, int kind, bool iota,
                  bool desc, const void** rkeys, const void** rpay, cudaStream_t s) {
    return sort_impl_items<U, P, HAS_PAY, 16>(src, bufA, bufB, psrc, pA, pB, n, kind, iota, desc, rkeys, rpay, s);
}

template <typename U>
int32_t sort_dispatch_pay(const void* src, void* bufA, void* bufB, int pay_bytes, const void* psrc, void* pA, void* pB,
                          int64_t n, int kind, bool iota, bool desc, const void** rkeys, const void** rpay,
                          cudaStream_t s) {
    if (pay_bytes == 0)
        return sort_impl<U, uint32_t, false>((const U*)src, (U*)bufA, (U*)bufB, nullptr, nullptr, nullptr, n, kind,
                                             false, desc, rkeys, rpay, s);
    if (pay_bytes == 4)
        return sort_impl<U, uint32_t, true>((const U*)src, (U*)bufA, (U*)bufB, (const uint32_t*)psrc, (uint32_t*)pA,
                                            (uint32_t*)pB, n, kind, iota, desc, rkeys, rpay, s);
    if (pay_bytes == 8)
        return sort_impl<U, uint64_t, true>((const U*)src, (U*)bufA, (U*)bufB, (const uint64_t*)psrc, (uint64_t*)pA,
                                            (uint64_t*)pB, n, kind, iota, desc, rkeys, rpay, s);
    set_error("radix sort: payload size %d", pay_bytes);
    return SDCB200_ERR_ARG;
}

}  // namespace

// src is only read; bufA / bufB (and pA / pB) are the ping-pong outputs.  *rkeys / *rpay point at
// whichever buffer holds the sorted rows afterwards (src itself when no pass had to run).
int32_t radix_sort_core(int dtype, const void* src, void* bufA, void* bufB, int pay_bytes, const void* psrc, void* pA,
                        void* pB, int64_t n, bool iota, bool desc, const void** rkeys, const void** rpay,
                        cudaStream_t s) {
    const int kind = dtype_kind(dtype);
    switch (dtype_size(dtype)) {
        case 1: return sort_dispatch_pay<uint8_t>(src, bufA, bufB, pay_bytes, psrc, pA, pB, n, kind, iota, desc, rkeys, rpay, s);
        case 2: return sort_dispatch_pay<uint16_t>(src, bufA, bufB, pay_bytes, psrc, pA, pB, n, kind, iota, desc, rkeys, rpay, s);
        case 4: return sort_dispatch_pay<uint32_t>(src, bufA, bufB, pay_bytes, psrc, pA, pB, n, kind, iota, desc, rkeys, rpay, s);
        case 8: return sort_dispatch_pay<uint64_t>(src, bufA, bufB, pay_bytes, psrc, pA, pB, n, kind, iota, desc, rkeys, rpay, s);
    }
    set_error("radix sort: bad dtype %d", dtype);
    return SDCB200_ERR_ARG;
}

// device stable argsort -> int64 row numbers
int32_t argsort_device(int dtype, const void* keys, int64_t n, bool ascending, int64_t* index_out, cudaStream_t s) {
    const int ks = dtype_size(dtype);
    SDC_ARG_CHECK(ks > 0, "dtype");
    if (n == 0) return SDCB200_OK;
    Scratch ka, kb, pa, pb;
    SDC_TRY(ka.alloc((size_t)n * ks, s));
    SDC_TRY(kb.alloc((size_t)n * ks, s));
    const void *rk, *rp;
    if (n <= 0xffffffffll) {
        SDC_TRY(pa.alloc((size_t)n * 4, s));
        SDC_TRY(pb.alloc((size_t)n * 4, s));
        SDC_TRY(radix_sort_core(dtype, keys, ka.p, kb.p, 4, nullptr, pa.p, pb.p, n, true, !ascending, &rk, &rp, s));
        widen_u32_kernel<<<sm_count() * 8, 256, 0, s>>>((const uint32_t*)rp, index_out, n);
        SDC_CHECK_LAUNCH();
    } else {
        SDC_TRY(pa.alloc((size_t)n * 8, s));
        SDC_TRY(radix_sort_core(dtype, keys, ka.p, kb.p, 8, nullptr, pa.p, index_out, n, true, !ascending, &rk, &rp, s));
        if (rp != index_out)
            SDC_CUDA_TRY(cudaMemcpyAsync(index_out, rp, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    }
    return SDCB200_OK;
}

}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

// parallel_sort_<t> / parallel_stable_sort_<t> (sdc/native/sort.cpp:119-121, stable_sort.cpp:308-310): in place.
int32_t sdcb200_sort(int32_t dtype, void* keys, int64_t n, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int ks = dtype_size(dtype);
    SDC_ARG_CHECK(ks > 0 && n >= 0, "dtype / n");
    if (n <= 1) return SDCB200_OK;
    SDC_ARG_CHECK(keys != nullptr, "null keys");
    Scratch alt;
    SDC_TRY(alt.alloc((size_t)n * ks, s));
    const void *rk, *rp;
    SDC_TRY(radix_sort_core(dtype, keys, alt.p, keys, 0, nullptr, nullptr, nullptr, n, false, false, &rk, &rp, s));
    if (rk != keys) SDC_CUDA_TRY(cudaMemcpyAsync(keys, rk, (size_t)n * ks, cudaMemcpyDeviceToDevice, s));
    return SDCB200_OK;
}

// parallel_stable_argsort_u64<t> (sdc/native/stable_sort.cpp:283-293): index_out[i] = row of the i-th smallest.
int32_t sdcb200_argsort(int32_t dtype, const void* keys, int64_t n, int32_t ascending, int64_t* index_out, void* stream) {
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(keys != nullptr && index_out != nullptr, "null pointer");
    return argsort_device(dtype, keys, n, ascending != 0, index_out, (cudaStream_t)stream);
}

// numpy_like.take (sdc/functions/numpy_like.py:1359-1388): dst[i] = src[idx[i]] for fixed-size rows.
int32_t sdcb200_gather(int32_t elem_size, const void* src, const int64_t* idx, int64_t n, void* dst, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    const int grid = sm_count() * 8;
    switch (elem_size) {
        case 1: gather_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)src, idx, (uint8_t*)dst, n); break;
        case 2: gather_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)src, idx, (uint16_t*)dst, n); break;
        case 4: gather_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)src, idx, (uint32_t*)dst, n); break;
        case 8: gather_kernel<uint64_t><<<grid, 256, 0, s>>>((const uint64_t*)src, idx, (uint64_t*)dst, n); break;
        default: set_error("gather: element size %d", elem_size); return SDCB200_ERR_ARG;
    }
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

int32_t sdcb200_host_sort(int32_t dtype, void* keys, int64_t n) {
    const int ks = dtype_size(dtype);
    SDC_ARG_CHECK(ks > 0 && n >= 0, "dtype / n");
    if (n <= 1) return SDCB200_OK;
    cudaStream_t s = nullptr;
    Scratch d;
    SDC_TRY(d.alloc((size_t)n * ks, s));
    SDC_CUDA_TRY(cudaMemcpyAsync(d.p, keys, (size_t)n * ks, cudaMemcpyHostToDevice, s));
    SDC_TRY(sdcb200_sort(dtype, d.p, n, s));
    SDC_CUDA_TRY(cudaMemcpyAsync(keys, d.p, (size_t)n * ks, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    return SDCB200_OK;
}

int32_t sdcb200_host_argsort(int32_t dtype, const void* keys, int64_t n, int32_t ascending, int64_t* index_out) {
    const int ks = dtype_size(dtype);
    SDC_ARG_CHECK(ks > 0 && n >= 0, "dtype / n");
    if (n == 0) return SDCB200_OK;
    cudaStream_t s = nullptr;
    Scratch d, di;
    SDC_TRY(d.alloc((size_t)n * ks, s));
    SDC_TRY(di.alloc((size_t)n * 8, s));
    SDC_CUDA_TRY(cudaMemcpyAsync(d.p, keys, (size_t)n * ks, cudaMemcpyHostToDevice, s));
    SDC_TRY(argsort_device(dtype, d.p, n, ascending != 0, di.as<int64_t>(), s));
    SDC_CUDA_TRY(cudaMemcpyAsync(index_out, di.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    return SDCB200_OK;
}

// ---- the reference's own symbol names (sdc/native/module.cpp:31-92), host buffers, void return ----
#define SDC_REF_SORT(T, CODE)                                                                              \
    void parallel_sort_##T(void* begin, uint64_t len) { sdcb200_host_sort(CODE, begin, (int64_t)len); }    \
    void parallel_stable_sort_##T(void* begin, uint64_t len) { sdcb200_host_sort(CODE, begin, (int64_t)len); } \
    void parallel_argsort_u64##T(void* index, void* begin, uint64_t len, uint8_t ascending) {              \
        sdcb200_host_argsort(CODE, begin, (int64_t)len, ascending, (int64_t*)index);                       \
    }                                                                                                      \
    void parallel_stable_argsort_u64##T(void* index, void* begin, uint64_t len, uint8_t ascending) {       \
        sdcb200_host_argsort(CODE, begin, (int64_t)len, ascending, (int64_t*)index);                       \
    }
SDC_REF_SORT(i8, SDCB200_I8)
SDC_REF_SORT(u8, SDCB200_U8)
SDC_REF_SORT(i16, SDCB200_I16)
SDC_REF_SORT(u16, SDCB200_U16)
SDC_REF_SORT(i32, SDCB200_I32)
SDC_REF_SORT(u32, SDCB200_U32)
SDC_REF_SORT(i64, SDCB200_I64)
SDC_REF_SORT(u64, SDCB200_U64)
SDC_REF_SORT(f32, SDCB200_F32)
SDC_REF_SORT(f64, SDCB200_F64)
#undef SDC_REF_SORT

// the reference sizes a TBB arena here (sdc/native/module.cpp:94-97); the GPU backend has no host pool
void set_number_of_threads(uint64_t) {}

}  // extern "C"
