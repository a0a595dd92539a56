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

// ------------------------------------------------------------------ partial tables of a low-cardinality groupby
// Every rank's partial result (one key row + one row per partial aggregate and column) travels as ONE int64 matrix
// [nrows][cap + 1] -- row r holds the ng values of source r zero-padded to cap, element [0][cap] the group count -- in a
// single all-gather; the gathered [P][nrows][cap + 1] block is then reduced over the ranks in ONE launch, which also
// decides whether that was legal (every rank found the same groups in the same order).
namespace sdcb200 { namespace {
constexpr int kPackMaxRows = 1 + 4 * 32;
struct PackArgs {
    const unsigned long long* src[kPackMaxRows];
    int nrows;
};
__global__ void pack_partials_kernel(PackArgs a, long long ng, long long cap, unsigned long long* __restrict__ dst) {
    const long long stride = cap + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)a.nrows * stride;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / stride, j = i - r * stride;
        unsigned long long v = 0ull;
        if (j < ng && ng <= cap) v = a.src[r][j];
        if (r == 0 && j == cap) v = (unsigned long long)ng;
        dst[i] = v;
    }
}

// ops per row of the matrix
enum { PART_KEY = 0, PART_SUM_F64 = 1, PART_SUM_I64 = 2, PART_MIN_F64 = 3, PART_MAX_F64 = 4, PART_MIN_I64 = 5, PART_MAX_I64 = 6 };
struct CombineArgs {
    int ops[kPackMaxRows];
    int nrows, P;
};
// flag[0] = 1 when some rank's key row or group count differs from rank 0's (zeroed by the caller); flag[1] = rank 0's count
__global__ void combine_partials_kernel(CombineArgs a, long long cap, const unsigned long long* __restrict__ all,
                                        unsigned long long* __restrict__ out, unsigned long long* __restrict__ flag) {
    const long long stride = cap + 1;
    const long long per_rank = (long long)a.nrows * stride;
    const long long ng0 = (long long)all[cap];
    if (blockIdx.x == 0 && threadIdx.x == 0) flag[1] = (unsigned long long)ng0;
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_rank; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / stride, j = i - r * stride;
        const int op = a.ops[r];
        unsigned long long acc = all[i];
        if (op == PART_KEY) {
            for (int q = 1; q < a.P; ++q) bad |= all[q * per_rank + i] != acc;      // keys and, at j == cap, the counts
            out[i] = acc;
            continue;
        }
        if (j >= ng0 || j >= cap) { out[i] = 0ull; continue; }
        for (int q = 1; q < a.P; ++q) {
            const unsigned long long v = all[q * per_rank + i];
            switch (op) {
                case PART_SUM_F64: acc = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)acc) + __longlong_as_double((long long)v)); break;
                case PART_SUM_I64: acc += v; break;
                case PART_MIN_F64:
                case PART_MAX_F64: {
                    // a partial is NaN when the rank saw no value of the group: skipna, NaN only when every rank's is
                    const double x = __longlong_as_double((long long)acc), y = __longlong_as_double((long long)v);
                    const double m = x != x ? y : (y != y ? x : (op == PART_MIN_F64 ? (y < x ? y : x) : (y > x ? y : x)));
                    acc = (unsigned long long)__double_as_longlong(m);
                    break;
                }
                case PART_MIN_I64: acc = (long long)v < (long long)acc ? v : acc; break;
                default: acc = (long long)v > (long long)acc ? v : acc; break;
            }
        }
        out[i] = acc;
    }
    if (bad) atomicExch(flag, 1ull);
}
}}

extern "C" int32_t sdcb200_pack_partials(const void* const* rows, int32_t nrows, int64_t ng, int64_t cap, int64_t* dst, void* stream) {
    SDC_ARG_CHECK(nrows >= 1 && nrows <= kPackMaxRows && ng >= 0 && cap >= 1 && dst != nullptr, "nrows / ng / cap");
    PackArgs a;
    a.nrows = nrows;
    for (int r = 0; r < nrows; ++r) a.src[r] = reinterpret_cast<const unsigned long long*>(rows[r]);
    const int64_t total = (int64_t)nrows * (cap + 1);
    int64_t grid = ceil_div(total, 256);
    if (grid > (int64_t)sm_count() * 4) grid = (int64_t)sm_count() * 4;
    pack_partials_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(a, ng, cap, reinterpret_cast<unsigned long long*>(dst));
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

extern "C" int32_t sdcb200_combine_partials(const int64_t* gathered, int32_t nranks, int32_t nrows, const int32_t* ops, int64_t cap,
                                            int64_t* out, int64_t* flag, void* stream) {
    SDC_ARG_CHECK(nranks >= 1 && nrows >= 1 && nrows <= kPackMaxRows && cap >= 1, "nranks / nrows / cap");
    SDC_ARG_CHECK(gathered != nullptr && ops != nullptr && out != nullptr && flag != nullptr, "null pointer");
    CombineArgs a;
    a.nrows = nrows;
    a.P = nranks;
    for (int r = 0; r < nrows; ++r) {
        SDC_ARG_CHECK(ops[r] >= PART_KEY && ops[r] <= PART_MAX_I64, "row op");
        a.ops[r] = ops[r];
    }
    cudaStream_t s = (cudaStream_t)stream;
    SDC_CUDA_TRY(cudaMemsetAsync(flag, 0, 16, s));
    const int64_t total = (int64_t)nrows * (cap + 1);
    int64_t grid = ceil_div(total, 256);
    if (grid > (int64_t)sm_count() * 4) grid = (int64_t)sm_count() * 4;
    combine_partials_kernel<<<(unsigned)grid, 256, 0, s>>>(a, cap, reinterpret_cast<const unsigned long long*>(gathered),
                                                          reinterpret_cast<unsigned long long*>(out),
                                                          reinterpret_cast<unsigned long long*>(flag));
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
// ROUTE_RANGE (sample sort, sort_values across GPUs): destination = the number of splitters <= the key's ordered image
// (order-preserving unsigned image, -0.0 == +0.0), reversed for a descending sort; NaN keys all go to one rank (the last
// one, or the first for na_position='first').  The key travels unchanged.
struct Route {
    int nparts, lg;
    int desc, nan_dest;                          // ROUTE_RANGE
    unsigned long long split[7];                 // ROUTE_RANGE: nparts - 1 ascending ordered images
};
enum { MODE_HASH = 0, MODE_MOD = 1, MODE_RANGE = 2 };
template <bool KEY_F64, int MODE>
__device__ __forceinline__ unsigned shuffle_dest(unsigned long long b, const Route& rt, unsigned long long& shipped) {
    if (MODE == MODE_RANGE) {
        shipped = b;
        unsigned long long c;
        if (KEY_F64) {
            const unsigned long long mag = b & 0x7fffffffffffffffull;
            if (mag > 0x7ff0000000000000ull) return (unsigned)rt.nan_dest;
            if (mag == 0) b = 0;
            c = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
        } else {
            c = b ^ 0x8000000000000000ull;
        }
        int d = 0;
#pragma unroll
        for (int i = 0; i < 7; ++i) d += (i < rt.nparts - 1 && c >= rt.split[i]) ? 1 : 0;
        return (unsigned)(rt.desc ? rt.nparts - 1 - d : d);
    }
    if (MODE == MODE_MOD) {
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

// Per-tile destination counts.  A thread counts its own rows in packed 8-bit fields (one per destination, <= 16 rows per
// thread), the fields are widened to 16 bits for the warp reduction (<= 512 per warp) and meet in shared counters: four
// instructions per row next to the load -- the earlier ballot-per-destination loop (8 votes per row) ran the count pass
// at 0.47 ms per 125 M rows, a third of the HBM rate.
template <bool KEY_F64, int MODE>
__global__ void __launch_bounds__(kShThreads) shuffle_count_kernel(const unsigned long long* __restrict__ keys, long long n,
                                                                   Route rt, unsigned* __restrict__ tile_counts,
                                                                   unsigned long long* __restrict__ counts) {
    __shared__ unsigned tot[kPeerMax];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < kPeerMax) tot[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kShTile + w * 32 * kShItems;
    unsigned long long k[kShItems];
#pragma unroll
    for (int i = 0; i < kShItems; ++i) {
        const long long r = base + i * 32 + lane;
        k[i] = r < n ? keys[r] : 0ull;
    }
    unsigned long long packed = 0ull;
#pragma unroll
    for (int i = 0; i < kShItems; ++i) {
        const long long r = base + i * 32 + lane;
        unsigned long long unused;
        const unsigned d = shuffle_dest<KEY_F64, MODE>(k[i], rt, unused);
        if (r < n) packed += 1ull << (8 * d);
    }
    unsigned long long even = packed & 0x00ff00ff00ff00ffull, odd = (packed >> 8) & 0x00ff00ff00ff00ffull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        even += __shfl_xor_sync(0xffffffffu, even, o);
        odd += __shfl_xor_sync(0xffffffffu, odd, o);
    }
    if (lane < kPeerMax) {
        const unsigned long long src = (lane & 1) ? odd : even;
        const unsigned c = (unsigned)(src >> (16 * (lane >> 1))) & 0xffffu;
        if (c) atomicAdd(&tot[lane], c);
    }
    __syncthreads();
    if (threadIdx.x < kPeerMax) {
        const unsigned t = tot[threadIdx.x];
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

// Partition-and-send.  512 threads x 8 rows (the same 4096-row tile as the count pass):
//   * the key loads and the first payload column's loads are issued together at the top, every further column's loads
//     while the previous one is being copied out -- the earlier kernel loaded a column only when it staged it and sat
//     on the long scoreboard for 44 % of its samples;
//   * a row's peers (rows of the same destination in its warp instruction) come from one vote per destination BIT
//     (three for eight ranks) instead of one per destination; the warp's running count per destination lives in shared
//     memory and is advanced by the lowest peer lane, so ranks are stable in row order;
//   * the staged tile leaves destination by destination (a run is one contiguous range of that rank's arena), with no
//     per-row search for the destination.
constexpr int kSdThreads = 512;
constexpr int kSdItems = kShTile / kSdThreads;      // 8
constexpr int kSdWarps = kSdThreads / 32;           // 16

constexpr int kSdStage = kShTile + 2 * kPeerMax;     // a run may start one slot late (see below)

template <bool KEY_F64, int MODE>
__global__ void __launch_bounds__(kSdThreads, 2) shuffle_send_kernel(SendArgs a, long long n, const unsigned* __restrict__ tile_off,
                                                                  int bulk) {
    // two staging buffers: column c + 1 is staged while the bulk stores of column c still read the other one
    extern __shared__ __align__(128) unsigned long long stage_all[];
    __shared__ unsigned wcnt[kSdWarps][kPeerMax];      // running count of (warp, destination); the warp total at the end
    __shared__ unsigned wbase[kSdWarps][kPeerMax];
    __shared__ unsigned stage_off[kPeerMax + 1];       // first staged slot of destination d
    __shared__ unsigned stage_cnt[kPeerMax];
    __shared__ long long run_dst[kPeerMax];            // arena row of the first staged row of destination d
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long tile_base = (long long)blockIdx.x * kShTile;
    const long long base = tile_base + w * 32 * kSdItems;
    const unsigned lt = (1u << lane) - 1u;
    unsigned long long cur[kSdItems], nxt[kSdItems];
    unsigned short slot[kSdItems];
    unsigned char dd[kSdItems];
#pragma unroll
    for (int i = 0; i < kSdItems; ++i) {
        const long long r = base + i * 32 + lane;
        cur[i] = r < n ? a.cols[0][r] : 0ull;
    }
    if (a.ncols > 1 && a.iota_col != 1) {
        const unsigned long long* col = a.cols[1];
#pragma unroll
        for (int i = 0; i < kSdItems; ++i) {
            const long long r = base + i * 32 + lane;
            nxt[i] = r < n ? col[r] : 0ull;
        }
    }
    if (lane < kPeerMax) wcnt[w][lane] = 0;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kSdItems; ++i) {
        const long long r = base + i * 32 + lane;
        const bool valid = r < n;
        unsigned long long shipped = 0ull;
        const unsigned d = shuffle_dest<KEY_F64, MODE>(cur[i], a.rt, shipped) & (kPeerMax - 1);
        cur[i] = shipped;
        dd[i] = valid ? (unsigned char)d : (unsigned char)0xff;
        unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const bool bit = (d >> b) & 1u;
            const unsigned v = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? v : ~v;
        }
        const unsigned prior = wcnt[w][d];
        __syncwarp();
        if (valid && (peers & lt) == 0u) wcnt[w][d] = prior + __popc(peers);      // the lowest peer lane
        __syncwarp();
        slot[i] = (unsigned short)(prior + __popc(peers & lt));
    }
    __syncthreads();
    if (tid < kSdWarps * kPeerMax) {
        const int ww = tid / kPeerMax, d = tid % kPeerMax;
        unsigned b = 0;
        for (int x = 0; x < ww; ++x) b += wcnt[x][d];
        wbase[ww][d] = b;
    }
    if (tid == 0) {
        unsigned acc = 0;
        for (int d = 0; d < kPeerMax; ++d) {
            unsigned t = 0;
            for (int x = 0; x < kSdWarps; ++x) t += wcnt[x][d];
            const long long dst = d < a.rt.nparts ? a.dst_off[d] + (long long)tile_off[(size_t)blockIdx.x * kPeerMax + d] : 0;
            // bulk stores want source and destination 16-byte aligned together: a run starts on a staged slot of the
            // same parity as its first arena row (the arena's row capacity is even, so every column agrees)
            if (bulk && ((acc ^ (unsigned)dst) & 1u)) ++acc;
            stage_off[d] = acc;
            stage_cnt[d] = t;
            run_dst[d] = dst;
            acc += t;
        }
        stage_off[kPeerMax] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSdItems; ++i)
        if (dd[i] != 0xff) slot[i] = (unsigned short)(stage_off[dd[i]] + wbase[w][dd[i]] + slot[i]);
    for (int c = 0; c < a.ncols; ++c) {
        unsigned long long* stage = stage_all + (size_t)(c & 1) * kSdStage;
        if (bulk && c >= 2) {
            if (lane == 0 && w < kPeerMax) tma_store_wait_read();      // column c - 2 has left this buffer
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < kSdItems; ++i)
            if (dd[i] != 0xff) stage[slot[i]] = cur[i];
        // the next column: already in registers (column 1), generated, or loaded now -- in flight during the copy-out
        if (c + 1 < a.ncols) {
            if (c + 1 == a.iota_col) {
#pragma unroll
                for (int i = 0; i < kSdItems; ++i) cur[i] = (unsigned long long)(a.iota_base + base + i * 32 + lane);
            } else if (c == 0) {
#pragma unroll
                for (int i = 0; i < kSdItems; ++i) cur[i] = nxt[i];
            } else {
                const unsigned long long* col = a.cols[c + 1];
#pragma unroll
                for (int i = 0; i < kSdItems; ++i) {
                    const long long r = base + i * 32 + lane;
                    cur[i] = r < n ? col[r] : 0ull;
                }
            }
        }
        if (bulk) {
            // ONE bulk (TMA) store per destination run: the copy engine moves each ~4 KB run to the peer's arena over
            // NVLink in large packets; an odd first / last row goes by a plain store
            fence_proxy_async();
            __syncthreads();
            if (lane == 0 && w < kPeerMax) {
                const int d = w;
                const unsigned cnt = stage_cnt[d];
                if (cnt) {
                    const unsigned lo = stage_off[d];
                    unsigned long long* dst = a.peer[d] + ((long long)c * a.cap + run_dst[d]);
                    const unsigned head = (unsigned)(run_dst[d] & 1);          // == lo & 1
                    if (head) dst[0] = stage[lo];
                    const unsigned body = (cnt - head) & ~1u;
                    if (body) {
                        tma_store_1d(dst + head, stage + lo + head, body * 8u);
                        tma_store_commit();
                    }
                    if (head + body < cnt) dst[cnt - 1] = stage[lo + cnt - 1];
                }
            }
        } else {
            __syncthreads();
#pragma unroll
            for (int d = 0; d < kPeerMax; ++d) {
                const int lo = (int)stage_off[d], hi = lo + (int)stage_cnt[d];
                unsigned long long* dst = a.peer[d] + ((long long)c * a.cap + run_dst[d] - lo);
                for (int q = lo + tid; q < hi; q += kSdThreads) dst[q] = stage[q];
            }
            __syncthreads();
        }
    }
    if (bulk && lane == 0 && w < kPeerMax) tma_store_wait_all();
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

namespace sdcb200 { namespace {
int32_t make_range_route(int32_t key_dtype, int32_t nparts, const uint64_t* splitters, int32_t descending, int32_t nan_dest, Route* rt) {
    SDC_ARG_CHECK(nparts >= 1 && nparts <= kPeerMax && (nparts == 1 || splitters != nullptr), "nparts / splitters");
    SDC_ARG_CHECK(nan_dest >= 0 && nan_dest < nparts, "nan_dest");
    memset(rt, 0, sizeof(*rt));
    rt->nparts = nparts;
    rt->lg = -1;
    rt->desc = descending ? 1 : 0;
    rt->nan_dest = nan_dest;
    for (int i = 0; i + 1 < nparts; ++i) {
        rt->split[i] = splitters[i];
        SDC_ARG_CHECK(i == 0 || splitters[i] >= splitters[i - 1], "splitters must ascend");
    }
    (void)key_dtype;
    return SDCB200_OK;
}

int32_t plan_impl(int32_t key_dtype, int mode, const Route& rt, const void* keys, int64_t n, int32_t nparts, uint32_t* tile_off,
                  int64_t* counts_out, cudaStream_t s) {
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(n >= 0 && n < 0xffffffffll && nparts >= 1 && nparts <= kPeerMax, "n (< 2^32) / nparts (<= 8)");
    SDC_ARG_CHECK(counts_out != nullptr && (n == 0 || (keys && tile_off)), "null pointer");
    SDC_CUDA_TRY(cudaMemsetAsync(counts_out, 0, (size_t)nparts * 8, s));
    if (n == 0) return SDCB200_OK;
    const int64_t ntiles = ceil_div(n, kShTile);
    const unsigned long long* k = reinterpret_cast<const unsigned long long*>(keys);
    Scratch cnt8;                                   // the kernels always keep kPeerMax counters
    SDC_TRY(cnt8.alloc(kPeerMax * 8, s));
    SDC_CUDA_TRY(cudaMemsetAsync(cnt8.p, 0, kPeerMax * 8, s));
    unsigned long long* c8 = cnt8.as<unsigned long long>();
    const bool f = key_dtype == SDCB200_F64;
    if (mode == MODE_MOD) shuffle_count_kernel<false, MODE_MOD><<<(unsigned)ntiles, kShThreads, 0, s>>>(k, n, rt, tile_off, c8);
    else if (mode == MODE_RANGE && f) shuffle_count_kernel<true, MODE_RANGE><<<(unsigned)ntiles, kShThreads, 0, s>>>(k, n, rt, tile_off, c8);
    else if (mode == MODE_RANGE) shuffle_count_kernel<false, MODE_RANGE><<<(unsigned)ntiles, kShThreads, 0, s>>>(k, n, rt, tile_off, c8);
    else if (f) shuffle_count_kernel<true, MODE_HASH><<<(unsigned)ntiles, kShThreads, 0, s>>>(k, n, rt, tile_off, c8);
    else shuffle_count_kernel<false, MODE_HASH><<<(unsigned)ntiles, kShThreads, 0, s>>>(k, n, rt, tile_off, c8);
    SDC_CHECK_LAUNCH();
    shuffle_tile_scan<<<kPeerMax, 1024, 0, s>>>(tile_off, ntiles);
    SDC_CHECK_LAUNCH();
    SDC_CUDA_TRY(cudaMemcpyAsync(counts_out, cnt8.p, (size_t)nparts * 8, cudaMemcpyDeviceToDevice, s));
    return SDCB200_OK;
}
}}

extern "C" int32_t sdcb200_shuffle_plan(int32_t key_dtype, int32_t route, const void* keys, int64_t n, int32_t nparts,
                                        uint32_t* tile_off, int64_t* counts_out, void* stream) {
    SDC_ARG_CHECK(nparts >= 1 && nparts <= kPeerMax, "nparts (<= 8)");
    Route rt;
    memset(&rt, 0, sizeof(rt));
    SDC_TRY(make_route(key_dtype, route, nparts, &rt));
    return plan_impl(key_dtype, route == SDCB200_ROUTE_MOD ? MODE_MOD : MODE_HASH, rt, keys, n, nparts, tile_off, counts_out, (cudaStream_t)stream);
}

extern "C" int32_t sdcb200_shuffle_plan_range(int32_t key_dtype, const uint64_t* splitters, int32_t descending, int32_t nan_dest,
                                              const void* keys, int64_t n, int32_t nparts, uint32_t* tile_off, int64_t* counts_out,
                                              void* stream) {
    Route rt;
    SDC_TRY(make_range_route(key_dtype, nparts, splitters, descending, nan_dest, &rt));
    return plan_impl(key_dtype, MODE_RANGE, rt, keys, n, nparts, tile_off, counts_out, (cudaStream_t)stream);
}

namespace sdcb200 { namespace {
int32_t send_impl(int32_t key_dtype, int mode, const Route& rt, int32_t ncols, const void* const* cols, int32_t iota_col,
                  int64_t iota_base, int64_t n, int32_t nparts, const uint32_t* tile_off, const int64_t* dst_off,
                  void* const* peer_bases, int64_t arena_rows, void* stream) {
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(ncols >= 1 && ncols <= kMaxCols, "ncols");
    SDC_ARG_CHECK(iota_col == -1 || (iota_col >= 1 && iota_col < ncols), "iota_col must be -1 or a payload column");
    SDC_ARG_CHECK(nparts >= 1 && nparts <= kPeerMax, "nparts (<= 8 GPUs of one node)");
    SDC_ARG_CHECK(n >= 0 && dst_off && peer_bases && cols, "null pointer");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(tile_off != nullptr, "null tile offsets");
    SendArgs a;
    memset(&a, 0, sizeof(a));
    a.rt = rt;
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
    // bulk (TMA) stores need 16-byte aligned arenas with an even row capacity (PeerShuffle allocates them so);
    // SDCB200_SHUFFLE_BULK=0 keeps the plain stores (A/B measurements)
    static const int bulk_env = [] { const char* e = getenv("SDCB200_SHUFFLE_BULK"); return e ? atoi(e) : 1; }();
    int bulk = bulk_env && (arena_rows & 1) == 0;
    for (int d = 0; d < nparts; ++d) if (reinterpret_cast<uintptr_t>(peer_bases[d]) & 15) bulk = 0;
    const size_t smem = (size_t)2 * kSdStage * 8;
    const bool f = key_dtype == SDCB200_F64;
    auto kern = mode == MODE_MOD ? shuffle_send_kernel<false, MODE_MOD>
                : mode == MODE_RANGE ? (f ? shuffle_send_kernel<true, MODE_RANGE> : shuffle_send_kernel<false, MODE_RANGE>)
                : (f ? shuffle_send_kernel<true, MODE_HASH> : shuffle_send_kernel<false, MODE_HASH>);
    SDC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
    kern<<<(unsigned)ntiles, kSdThreads, smem, s>>>(a, n, tile_off, bulk);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}
}}

extern "C" int32_t sdcb200_shuffle_send(int32_t key_dtype, int32_t route, int32_t ncols, const void* const* cols,
                                        int32_t iota_col, int64_t iota_base, int64_t n, int32_t nparts,
                                        const uint32_t* tile_off, const int64_t* dst_off, void* const* peer_bases,
                                        int64_t arena_rows, void* stream) {
    SDC_ARG_CHECK(nparts >= 1 && nparts <= kPeerMax, "nparts (<= 8 GPUs of one node)");
    Route rt;
    memset(&rt, 0, sizeof(rt));
    SDC_TRY(make_route(key_dtype, route, nparts, &rt));
    return send_impl(key_dtype, route == SDCB200_ROUTE_MOD ? MODE_MOD : MODE_HASH, rt, ncols, cols, iota_col, iota_base, n, nparts,
                     tile_off, dst_off, peer_bases, arena_rows, stream);
}

extern "C" int32_t sdcb200_shuffle_send_range(int32_t key_dtype, const uint64_t* splitters, int32_t descending, int32_t nan_dest,
                                              int32_t ncols, const void* const* cols, int32_t iota_col, int64_t iota_base, int64_t n,
                                              int32_t nparts, const uint32_t* tile_off, const int64_t* dst_off,
                                              void* const* peer_bases, int64_t arena_rows, void* stream) {
    Route rt;
    SDC_TRY(make_range_route(key_dtype, nparts, splitters, descending, nan_dest, &rt));
    return send_impl(key_dtype, MODE_RANGE, rt, ncols, cols, iota_col, iota_base, n, nparts, tile_off, dst_off, peer_bases,
                     arena_rows, stream);
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
