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
#include <cstring>

#include "radix_sort.cuh"

namespace sdcb200 {
namespace {

constexpr int kPartThreads = 256;
constexpr int kMaxParts = 256;
constexpr int kMaxCols = 32;

__host__ __device__ inline unsigned long long pmix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// route 1 (SDCB200_ROUTE_MOD, int64 keys): destination = key mod nparts (Euclidean), and `shipped` receives
// floor(key / nparts) -- see Route / shuffle_dest below, which the peer-memory path shares
template <bool KEY_F64, bool MOD>
__global__ void __launch_bounds__(kPartThreads) partition_dest_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                                                      unsigned nparts, uint8_t* __restrict__ dest,
                                                                      unsigned long long* __restrict__ counts,
                                                                      unsigned long long* __restrict__ shipped) {
    __shared__ unsigned hist[kMaxParts];
    for (int i = threadIdx.x; i < kMaxParts; i += kPartThreads) hist[i] = 0;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)kPartThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kPartThreads) {
        unsigned long long b = keys[i];
        unsigned d;
        if (MOD) {
            const long long k = (long long)b;
            long long q = k / (long long)nparts, r = k - q * (long long)nparts;
            if (r < 0) { r += nparts; q -= 1; }
            d = (unsigned)r;
            if (shipped) shipped[i] = (unsigned long long)q;
        } else {
            if (shipped) shipped[i] = b;
            if (KEY_F64) {
                const unsigned long long mag = b & 0x7fffffffffffffffull;
                if (mag > 0x7ff0000000000000ull) b = 0x7ff8000000000000ull;
                else if (mag == 0) b = 0;
            }
            // the TOP bits choose the rank, so the low bits the local hash tables index with stay uniform
            d = (unsigned)(((pmix64(b) >> 32) * (unsigned long long)nparts) >> 32);
        }
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

extern "C" int32_t sdcb200_hash_partition(int32_t key_dtype, int32_t route, const void* keys, int64_t n, int32_t nparts,
                                          int64_t* perm_out, int64_t* counts_out, void* shipped_keys_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(route == SDCB200_ROUTE_HASH || (route == SDCB200_ROUTE_MOD && key_dtype == SDCB200_I64), "route (MOD needs int64 keys)");
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
    unsigned long long* sh = reinterpret_cast<unsigned long long*>(shipped_keys_out);
    if (route == SDCB200_ROUTE_MOD)
        partition_dest_kernel<false, true><<<(unsigned)grid, kPartThreads, 0, s>>>(k, n, (unsigned)nparts, dest.as<uint8_t>(), c, sh);
    else if (key_dtype == SDCB200_F64)
        partition_dest_kernel<true, false><<<(unsigned)grid, kPartThreads, 0, s>>>(k, n, (unsigned)nparts, dest.as<uint8_t>(), c, sh);
    else
        partition_dest_kernel<false, false><<<(unsigned)grid, kPartThreads, 0, s>>>(k, n, (unsigned)nparts, dest.as<uint8_t>(), c, sh);
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

// ------------------------------------------------------------------ peer-memory shuffle (NVLink P2P stores)
// One node, one process per GPU: every rank owns receive arenas (cudaMalloc, exported with CUDA IPC handles and
// opened by its peers).  Rows are stored straight into the destination ranks' arenas by the fused
// partition-and-send kernel below; there is no packed send buffer and no per-column all-to-all.
namespace sdcb200 { namespace {
constexpr int kPeerMax = 8;
}}

extern "C" int32_t sdcb200_ipc_alloc(void** p, int64_t bytes) {
    SDC_ARG_CHECK(p != nullptr && bytes > 0, "ipc_alloc");
    SDC_CUDA_TRY(cudaMalloc(p, (size_t)bytes));
    return SDCB200_OK;
}
extern "C" int32_t sdcb200_ipc_free(void* p) {
    if (p) SDC_CUDA_TRY(cudaFree(p));
    return SDCB200_OK;
}
extern "C" int32_t sdcb200_ipc_get_handle(void* p, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    SDC_ARG_CHECK(p != nullptr && handle64 != nullptr, "ipc_get_handle");
    cudaIpcMemHandle_t h;
    SDC_CUDA_TRY(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    return SDCB200_OK;
}
extern "C" int32_t sdcb200_ipc_open(const void* handle64, void** p) {
    SDC_ARG_CHECK(p != nullptr && handle64 != nullptr, "ipc_open");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SDC_CUDA_TRY(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
    return SDCB200_OK;
}
extern "C" int32_t sdcb200_ipc_close(void* p) {
    if (p) SDC_CUDA_TRY(cudaIpcCloseMemHandle(p));
    return SDCB200_OK;
}

// ------------------------------------------------------------------ fused partition + send
// No row permutation is ever materialised: tiles of 4096 rows, two kernels.
//   shuffle_count_kernel   destination of every row (same hash as sdcb200_hash_partition), per-tile counts per
//                          destination (ballots), total counts
//   shuffle_tile_scan      exclusive scan of the tile counts per destination -> tile_off[tile][d]
//   (host: all-gather of the totals -> this source's first row in every destination arena)
//   shuffle_send_kernel    recomputes the destinations, ranks the rows of each destination inside the tile in row
//                          order (ballot multisplit + warp prefix: stable), stages one column at a time in shared
//                          memory in destination order and copies every destination's run -- ~4 KB contiguous --
//                          to that rank's arena over NVLink.  Same row order in the arenas as the NCCL path.
namespace sdcb200 { namespace {
constexpr int kShThreads = 256;
constexpr int kShItems = 16;
constexpr int kShTile = kShThreads * kShItems;      // 4096 rows
constexpr int kShWarps = kShThreads / 32;

// Routing of a row.  ROUTE_HASH: destination = top bits of mix64(canonical key), the key travels unchanged.
// ROUTE_MOD (int64 keys): destination = key mod P (Euclidean) and the key that travels is floor(key / P), so a
// dense key range [lo, hi] arrives as the dense range [lo / P, hi / P] on every rank -- the local join / groupby
// keep their direct-addressed tables after the shuffle (hashing the keys destroys exactly that).  The owner maps
// a shipped key back with key = q * P + rank.  lg = log2(P) when P is a power of two, else -1.
struct Route {
    int nparts, lg;
};
template <bool KEY_F64, bool MOD>
__device__ __forceinline__ unsigned shuffle_dest(unsigned long long b, const Route& rt, unsigned long long& shipped) {
    if (MOD) {
        const long long k = (long long)b;
        if (rt.lg >= 0) {
            shipped = (unsigned long long)(k >> rt.lg);
            return (unsigned)(b & (unsigned long long)(rt.nparts - 1));
        }
        long long q = k / rt.nparts, r = k - q * rt.nparts;
        if (r < 0) { r += rt.nparts; q -= 1; }
        shipped = (unsigned long long)q;
        return (unsigned)r;
    }
    shipped = b;
    if (KEY_F64) {
        const unsigned long long mag = b & 0x7fffffffffffffffull;
        if (mag > 0x7ff0000000000000ull) b = 0x7ff8000000000000ull;
        else if (mag == 0) b = 0;
    }
    return (unsigned)(((pmix64(b) >> 32) * (unsigned long long)rt.nparts) >> 32);
}

template <bool KEY_F64, bool MOD>
__global__ void __launch_bounds__(kShThreads) shuffle_count_kernel(const unsigned long long* __restrict__ keys, long long n,
                                                                   Route rt, unsigned* __restrict__ tile_counts,
                                                                   unsigned long long* __restrict__ counts) {
    __shared__ unsigned wtot[kShWarps][kPeerMax];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * kShTile + w * 32 * kShItems;
    unsigned run[kPeerMax];
#pragma unroll
    for (int d = 0; d < kPeerMax; ++d) run[d] = 0;
#pragma unroll
    for (int i = 0; i < kShItems; ++i) {
        const long long r = base + i * 32 + lane;
        unsigned long long unused;
        const unsigned d = r < n ? shuffle_dest<KEY_F64, MOD>(keys[r], rt, unused) : 0xffu;
#pragma unroll
        for (int q = 0; q < kPeerMax; ++q) run[q] += __popc(__ballot_sync(0xffffffffu, d == (unsigned)q));
    }
#pragma unroll
    for (int q = 0; q < kPeerMax; ++q) if (lane == q) wtot[w][q] = run[q];
    __syncthreads();
    if (threadIdx.x < kPeerMax) {
        unsigned t = 0;
        for (int ww = 0; ww < kShWarps; ++ww) t += wtot[ww][threadIdx.x];
        tile_counts[(size_t)blockIdx.x * kPeerMax + threadIdx.x] = t;
        if (t) atomicAdd(&counts[threadIdx.x], (unsigned long long)t);
    }
}

// grid = kPeerMax CTAs of 1024 threads: CTA d scans column d of tile_counts in place (exclusive)
__global__ void __launch_bounds__(1024) shuffle_tile_scan(unsigned* __restrict__ tile_counts, long long ntiles) {
    __shared__ unsigned wsum[32];
    const int d = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long per = (ntiles + 1023) / 1024;
    const long long lo = (long long)threadIdx.x * per;
    const long long hi = lo + per < ntiles ? lo + per : ntiles;
    unsigned s = 0;
    for (long long t = lo; t < hi; ++t) s += tile_counts[t * kPeerMax + d];
    unsigned incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned basev = 0;
    for (int ww = 0; ww < warp; ++ww) basev += wsum[ww];
    unsigned run = basev + incl - s;
    for (long long t = lo; t < hi; ++t) {
        const unsigned v = tile_counts[t * kPeerMax + d];
        tile_counts[t * kPeerMax + d] = run;
        run += v;
    }
}

struct SendArgs {
    const unsigned long long* cols[kMaxCols];   // cols[0] = the key column
    unsigned long long* peer[kPeerMax];
    long long dst_off[kPeerMax];
    long long cap;
    long long iota_base;        // column `iota_col` is not read: its value for row r is iota_base + r (global row ids)
    int ncols, iota_col;
    Route rt;
};

template <bool KEY_F64, bool MOD>
__global__ void __launch_bounds__(kShThreads) shuffle_send_kernel(SendArgs a, long long n, const unsigned* __restrict__ tile_off) {
    __shared__ unsigned long long stage[kShTile];
    __shared__ unsigned wtot[kShWarps][kPeerMax];
    __shared__ unsigned wbase[kShWarps][kPeerMax];
    __shared__ unsigned stage_off[kPeerMax + 1];
    __shared__ long long run_dst[kPeerMax];            // global row of the first staged row of destination d
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long tile_base = (long long)blockIdx.x * kShTile;
    const long long base = tile_base + w * 32 * kShItems;
    const int tile_n = (int)(n - tile_base < kShTile ? n - tile_base : kShTile);
    const unsigned lt = (1u << lane) - 1u;
    unsigned long long key[kShItems];
    unsigned short slot[kShItems];       // rank inside (warp, destination) first, staged slot later
    unsigned char dd[kShItems];
    unsigned run[kPeerMax];
#pragma unroll
    for (int d = 0; d < kPeerMax; ++d) run[d] = 0;
#pragma unroll
    for (int i = 0; i < kShItems; ++i) {
        const long long r = base + i * 32 + lane;
        key[i] = r < n ? a.cols[0][r] : 0ull;
    }
#pragma unroll
    for (int i = 0; i < kShItems; ++i) {
        const long long r = base + i * 32 + lane;
        unsigned long long shipped = 0ull;
        const unsigned d = r < n ? shuffle_dest<KEY_F64, MOD>(key[i], a.rt, shipped) : 0xffu;
        key[i] = shipped;
        dd[i] = (unsigned char)d;
        unsigned mine = 0;
#pragma unroll
        for (int q = 0; q < kPeerMax; ++q) {
            const unsigned b = __ballot_sync(0xffffffffu, d == (unsigned)q);
            if (d == (unsigned)q) mine = run[q] + __popc(b & lt);
            run[q] += __popc(b);
        }
        slot[i] = (unsigned short)mine;
    }
#pragma unroll
    for (int q = 0; q < kPeerMax; ++q) if (lane == q) wtot[w][q] = run[q];
    __syncthreads();
    if (tid < kShWarps * kPeerMax) {
        const int ww = tid / kPeerMax, d = tid % kPeerMax;
        unsigned b = 0;
        for (int x = 0; x < ww; ++x) b += wtot[x][d];
        wbase[ww][d] = b;
    }
    if (tid == 0) {
        unsigned acc = 0;
        for (int d = 0; d < kPeerMax; ++d) {
            unsigned t = 0;
            for (int x = 0; x < kShWarps; ++x) t += wtot[x][d];
            stage_off[d] = acc;
            run_dst[d] = d < a.rt.nparts ? a.dst_off[d] + (long long)tile_off[(size_t)blockIdx.x * kPeerMax + d] : 0;
            acc += t;
        }
        stage_off[kPeerMax] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kShItems; ++i)
        if (dd[i] != 0xff) slot[i] = (unsigned short)(stage_off[dd[i]] + wbase[w][dd[i]] + slot[i]);
    for (int c = 0; c < a.ncols; ++c) {
        const unsigned long long* col = a.cols[c];
#pragma unroll
        for (int i = 0; i < kShItems; ++i) {
            const long long r = base + i * 32 + lane;
            if (dd[i] != 0xff) stage[slot[i]] = c == 0 ? key[i] : (c == a.iota_col ? (unsigned long long)(a.iota_base + r) : col[r]);
        }
        __syncthreads();
        for (int q = tid; q < tile_n; q += kShThreads) {
            int d = 0;
#pragma unroll
            for (int x = 1; x < kPeerMax; ++x) d += (q >= (int)stage_off[x]) ? 1 : 0;
            a.peer[d][(long long)c * a.cap + run_dst[d] + (q - (int)stage_off[d])] = stage[q];
        }
        __syncthreads();
    }
}
}}

namespace sdcb200 { namespace {
int32_t make_route(int32_t key_dtype, int32_t route, int32_t nparts, Route* rt) {
    SDC_ARG_CHECK(route == SDCB200_ROUTE_HASH || route == SDCB200_ROUTE_MOD, "route must be HASH or MOD");
    SDC_ARG_CHECK(route == SDCB200_ROUTE_HASH || key_dtype == SDCB200_I64, "ROUTE_MOD needs int64 keys");
    rt->nparts = nparts;
    rt->lg = -1;
    for (int b = 0; b < 8; ++b) if ((1 << b) == nparts) rt->lg = b;
    return SDCB200_OK;
}
}}

extern "C" int32_t sdcb200_shuffle_plan(int32_t key_dtype, int32_t route, const void* keys, int64_t n, int32_t nparts,
                                        uint32_t* tile_off, int64_t* counts_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(n >= 0 && n < 0xffffffffll && nparts >= 1 && nparts <= kPeerMax, "n (< 2^32) / nparts (<= 8)");
    SDC_ARG_CHECK(counts_out != nullptr && (n == 0 || (keys && tile_off)), "null pointer");
    Route rt;
    SDC_TRY(make_route(key_dtype, route, nparts, &rt));
    SDC_CUDA_TRY(cudaMemsetAsync(counts_out, 0, (size_t)nparts * 8, s));
    if (n == 0) return SDCB200_OK;
    const int64_t ntiles = ceil_div(n, kShTile);
    const unsigned long long* k = reinterpret_cast<const unsigned long long*>(keys);
    Scratch cnt8;                                   // the kernels always keep kPeerMax counters
    SDC_TRY(cnt8.alloc(kPeerMax * 8, s));
    SDC_CUDA_TRY(cudaMemsetAsync(cnt8.p, 0, kPeerMax * 8, s));
    if (route == SDCB200_ROUTE_MOD)
        shuffle_count_kernel<false, true><<<(unsigned)ntiles, kShThreads, 0, s>>>(k, n, rt, tile_off, cnt8.as<unsigned long long>());
    else if (key_dtype == SDCB200_F64)
        shuffle_count_kernel<true, false><<<(unsigned)ntiles, kShThreads, 0, s>>>(k, n, rt, tile_off, cnt8.as<unsigned long long>());
    else
        shuffle_count_kernel<false, false><<<(unsigned)ntiles, kShThreads, 0, s>>>(k, n, rt, tile_off, cnt8.as<unsigned long long>());
    SDC_CHECK_LAUNCH();
    shuffle_tile_scan<<<kPeerMax, 1024, 0, s>>>(tile_off, ntiles);
    SDC_CHECK_LAUNCH();
    SDC_CUDA_TRY(cudaMemcpyAsync(counts_out, cnt8.p, (size_t)nparts * 8, cudaMemcpyDeviceToDevice, s));
    return SDCB200_OK;
}

extern "C" int32_t sdcb200_shuffle_send(int32_t key_dtype, int32_t route, int32_t ncols, const void* const* cols,
                                        int32_t iota_col, int64_t iota_base, int64_t n, int32_t nparts,
                                        const uint32_t* tile_off, const int64_t* dst_off, void* const* peer_bases,
                                        int64_t arena_rows, void* stream) {
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(ncols >= 1 && ncols <= kMaxCols, "ncols");
    SDC_ARG_CHECK(iota_col == -1 || (iota_col >= 1 && iota_col < ncols), "iota_col must be -1 or a payload column");
    SDC_ARG_CHECK(nparts >= 1 && nparts <= kPeerMax, "nparts (<= 8 GPUs of one node)");
    SDC_ARG_CHECK(n >= 0 && dst_off && peer_bases && cols, "null pointer");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(tile_off != nullptr, "null tile offsets");
    SendArgs a;
    memset(&a, 0, sizeof(a));
    SDC_TRY(make_route(key_dtype, route, nparts, &a.rt));
    for (int c = 0; c < ncols; ++c) {
        SDC_ARG_CHECK(c == iota_col || cols[c] != nullptr, "null column");
        a.cols[c] = reinterpret_cast<const unsigned long long*>(cols[c]);
    }
    for (int d = 0; d < nparts; ++d) {
        a.peer[d] = reinterpret_cast<unsigned long long*>(peer_bases[d]);
        a.dst_off[d] = dst_off[d];
        SDC_ARG_CHECK(dst_off[d] >= 0 && dst_off[d] <= arena_rows, "arena offset");
    }
    a.cap = arena_rows;
    a.ncols = ncols;
    a.iota_col = iota_col;
    a.iota_base = iota_base;
    const int64_t ntiles = ceil_div(n, kShTile);
    cudaStream_t s = (cudaStream_t)stream;
    if (route == SDCB200_ROUTE_MOD)
        shuffle_send_kernel<false, true><<<(unsigned)ntiles, kShThreads, 0, s>>>(a, n, tile_off);
    else if (key_dtype == SDCB200_F64)
        shuffle_send_kernel<true, false><<<(unsigned)ntiles, kShThreads, 0, s>>>(a, n, tile_off);
    else
        shuffle_send_kernel<false, false><<<(unsigned)ntiles, kShThreads, 0, s>>>(a, n, tile_off);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

// ------------------------------------------------------------------ small row-wise helpers of the distributed pass
namespace sdcb200 { namespace {
// x -> x * mul + add  (rows < 0 stay as they are when keep_neg: the -1 "no partner" marker of a left join)
__global__ void affine_i64_kernel(long long* __restrict__ x, int64_t n, long long mul, long long add, int keep_neg) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long v = x[i];
        if (!(keep_neg && v < 0)) x[i] = v * mul + add;
    }
}
// dst[i] = idx[i] < 0 ? -1 : src[idx[i]]: local row numbers of a join -> the global row ids that travelled with the rows
__global__ void gather_ids_kernel(const long long* __restrict__ src, const long long* __restrict__ idx, long long* __restrict__ dst,
                                  int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long j = __ldcs(idx + i);
        __stcs(dst + i, j < 0 ? -1ll : __ldg(src + j));
    }
}
}}

extern "C" int32_t sdcb200_affine_i64(int64_t* x, int64_t n, int64_t mul, int64_t add, int32_t keep_negative, void* stream) {
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(x != nullptr, "null column");
    int64_t grid = ceil_div(n, 256 * 4);
    if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
    affine_i64_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(x), n, mul, add, keep_negative);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

extern "C" int32_t sdcb200_gather_ids(const int64_t* src, const int64_t* idx, int64_t n, int64_t* dst, void* stream) {
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(src != nullptr && idx != nullptr && dst != nullptr, "null column");
    int64_t grid = ceil_div(n, 256 * 4);
    if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
    gather_ids_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(src),
                                                                         reinterpret_cast<const long long*>(idx),
                                                                         reinterpret_cast<long long*>(dst), n);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}
