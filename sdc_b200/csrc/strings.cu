// strings.cu -- str_arr columns (offsets + chars in device memory) as sort and groupby keys (K4, K6).
//
// Column layout = the reference's StringArray unchanged (sdc/str_arr_type.py:33-101,
// sdc/native/str_ext/sdc_str_arr.hpp:160-176): uint32 offsets[n + 1] (offsets[0] = 0,
// offsets[n] = total chars), uint8 chars[total] UTF-8 without terminators, validity bitmap
// LSB-first with bit = 1 meaning valid; NA strings are stored empty (:389-393).
//
//   sdcb200_str_argsort   replaces `stable_argsort` (sdc_str_arr.hpp:565-593, bound at
//       sdc/str_arr_ext.py:94-99): the reference copies every string into a std::string and runs a
//       SERIAL std::stable_sort of (string, idx) pairs with std::string `<` / `>` -- unsigned
//       byte-lexicographic order, a proper prefix first, ties in input order for both directions.
//       Here: LSD radix sort over the strings' 8-byte big-endian chunks, last chunk first, preceded by
//       one pass on the length (the least significant key: zero padding makes "ab" and "ab\0" equal
//       chunk-wise, the length puts the prefix first).  Every pass is the library's stable onesweep
//       sort of (u64 chunk key, u32 row), so constant digits are skipped and ties keep input order.
//   sdcb200_str_factorize  string keys -> dense group ids for the hash groupby: open-addressing table
//       of 16-byte slots {key word, length << 32 | representative row}; strings of up to 8 bytes ARE the
//       key word (no look at the representative), longer ones hash into it and a hit is confirmed by
//       comparing the bytes, so ids are exact.  gid = slot index in [0, cap], -1 for NA rows (None keys
//       are skipped, reference hpat_pandas_dataframe_functions.py:3088-3089).  The numeric groupby then
//       runs on the ids (key dtype SDCB200_GID: direct-addressed table).
//   sdcb200_str_gather     rows -> a new str_arr (lengths, exclusive scan, byte copy).
#include <cmath>

#include "radix_sort.cuh"
#include "scan.cuh"

namespace sdcb200 {
namespace {

constexpr int kStrThreads = 256;

__device__ __forceinline__ uint64_t chunk_key(const uint8_t* __restrict__ chars, uint32_t beg, uint32_t end, uint32_t chunk) {
    // big-endian bytes [8*chunk, 8*chunk + 8) of the string, zero padded
    const uint64_t p = (uint64_t)beg + 8ull * chunk;
    uint64_t k = 0;
    if (p >= end) return 0;
    const uint32_t avail = (uint32_t)(end - p < 8 ? end - p : 8);
#pragma unroll
    for (uint32_t b = 0; b < 8; ++b) {
        k <<= 8;
        if (b < avail) k |= chars[p + b];
    }
    return k;
}

// keys for one LSD pass: chunk < 0 -> the length, else the chunk's bytes; rows taken through `perm`
__global__ void __launch_bounds__(kStrThreads) str_pass_keys_kernel(const uint8_t* __restrict__ chars,
                                                                    const uint32_t* __restrict__ offsets,
                                                                    const uint32_t* __restrict__ perm, int64_t n, int chunk,
                                                                    uint64_t* __restrict__ keys) {
    for (int64_t i = blockIdx.x * (int64_t)kStrThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStrThreads) {
        const uint32_t r = perm ? perm[i] : (uint32_t)i;
        const uint32_t beg = offsets[r], end = offsets[r + 1];
        keys[i] = chunk < 0 ? (uint64_t)(end - beg) : chunk_key(chars, beg, end, (uint32_t)chunk);
    }
}

__global__ void __launch_bounds__(kStrThreads) str_maxlen_kernel(const uint32_t* __restrict__ offsets, int64_t n,
                                                                 unsigned* __restrict__ out /* [2]: max, min */) {
    unsigned mx = 0, mn = 0xffffffffu;
    for (int64_t i = blockIdx.x * (int64_t)kStrThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStrThreads) {
        const unsigned len = offsets[i + 1] - offsets[i];
        mx = len > mx ? len : mx;
        mn = len < mn ? len : mn;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned a = __shfl_xor_sync(0xffffffffu, mx, d), b = __shfl_xor_sync(0xffffffffu, mn, d);
        mx = a > mx ? a : mx;
        mn = b < mn ? b : mn;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, mx);
        atomicMin(out + 1, mn);
    }
}

__global__ void widen_rows_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in ? in[i] : (uint64_t)i;
}

inline int grid_for(int64_t n, int threads, int per_sm) {
    const int64_t want = ceil_div(n, threads);
    const int64_t cap = (int64_t)sm_count() * per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// ---- factorize ----
// Table of 16-byte slots {key word, len << 32 | group id}, linear probing, the key word claimed with one 64-bit
// atomicCAS; the second word is published right after and readers wait for it.  Group ids are DENSE, in order of
// insertion (the claimer's ticket from the group counter), so the aggregate that follows runs on ids in
// [0, ngroups) -- a direct-addressed table exactly as large as the result -- whatever the capacity of this table,
// which can therefore be roomy (load factor <= 1/8 for small group counts: a warp's 32 probe loops end after one
// or two steps instead of six to eight, and the kernel is issue-bound).  rep[id] = the inserting row.
//   len <= 8 : key word = the string's bytes (little-endian, zero padded) -- together with the length in
//              the second word that IS the string, so a match needs no look at the representative row:
//              one 16-byte slot load per probe, no dependent random reads (C4's 8-character keys)
//   len  > 8 : key word = 64-bit hash of the bytes; a (word, length) match is confirmed by comparing the bytes
//              with the representative row rep[id], so ids are exact
// A key word equal to the empty marker lives in the spare slot [cap].
constexpr uint64_t kStrEmpty = ~0ull;
constexpr uint64_t kStrPending = ~0ull;

// bytes [beg, beg + len) with len <= 8 as a little-endian word, zero padded
__device__ __forceinline__ uint64_t load_word(const uint8_t* __restrict__ chars, uint64_t beg, uint32_t len, uint64_t total) {
    if (len == 0) return 0ull;
    if ((beg & 7ull) == 0 && beg + 8 <= total) {
        const uint64_t w = *reinterpret_cast<const uint64_t*>(chars + beg);
        return len >= 8 ? w : (w & ((1ull << (8 * len)) - 1ull));
    }
    uint64_t w = 0;
    for (uint32_t b = 0; b < len; ++b) w |= (uint64_t)chars[beg + b] << (8 * b);
    return w;
}

__device__ __forceinline__ uint64_t smix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__device__ __forceinline__ uint64_t str_hash(const uint8_t* __restrict__ chars, uint64_t beg, uint32_t len, uint64_t total) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ len;
    for (uint32_t o = 0; o < len; o += 8) h = smix64(h ^ load_word(chars, beg + o, len - o < 8 ? len - o : 8, total));
    return h;
}

__device__ __forceinline__ bool str_equal(const uint8_t* __restrict__ chars, uint64_t a, uint64_t b, uint32_t len, uint64_t total) {
    for (uint32_t o = 0; o < len; o += 8) {
        const uint32_t m = len - o < 8 ? len - o : 8;
        if (load_word(chars, a + o, m, total) != load_word(chars, b + o, m, total)) return false;
    }
    return true;
}

__global__ void __launch_bounds__(kStrThreads) str_factorize_kernel(const uint8_t* __restrict__ chars,
                                                                    const uint32_t* __restrict__ offsets,
                                                                    const uint8_t* __restrict__ bitmap, int64_t n,
                                                                    ulonglong2* __restrict__ table, uint64_t cap,
                                                                    uint32_t* __restrict__ rep /* [limit + 2] */,
                                                                    int64_t* __restrict__ gid,
                                                                    unsigned long long* __restrict__ counters /* [0] groups, [1] overflow */,
                                                                    uint64_t limit) {
    const uint64_t total = offsets[n];
    const uint64_t stream_pol = l2_policy_evict_first();
    const int lane = threadIdx.x & 31;
    // the claimer of a slot: ticket from the group counter, representative row, then the slot's second word
    auto publish = [&](unsigned long long* second, uint32_t len, uint32_t row) -> int64_t {
        const unsigned long long id = atomicAdd(&counters[0], 1ull);
        if (id + 1 > limit) atomicExch(&counters[1], 1ull);
        if (id <= limit) rep[id] = row;
        __threadfence();                                   // rep[id] is visible before the id is
        *reinterpret_cast<volatile unsigned long long*>(second) = ((uint64_t)len << 32) | (uint64_t)(uint32_t)id;
        return (int64_t)id;
    };
    // warp-uniform loop (a warp owns 32 consecutive rows per step) and an explicit reconvergence before the
    // coalesced accesses: the probe loop's trip count differs per lane, and a warp that stays split issues its
    // loads and stores as several partial requests
    for (int64_t w0 = blockIdx.x * (int64_t)kStrThreads + (threadIdx.x & ~31); w0 < n; w0 += (int64_t)gridDim.x * kStrThreads) {
        __syncwarp();
        const int64_t i = w0 + lane;
        const bool in = i < n;
        const bool valid = in && !(bitmap && !((bitmap[i >> 3] >> (i & 7)) & 1));
        uint32_t beg = 0, len = 0;
        if (valid) { beg = offsets[i]; len = offsets[i + 1] - beg; }
        uint64_t kw = 0;
        if (valid) kw = len <= 8 ? load_word(chars, beg, len, total) : str_hash(chars, beg, len, total);
        if (len > 8 && kw == kStrEmpty) kw = kStrEmpty - 1;      // only the 8-byte string of 0xFF uses the spare slot
        int64_t out = valid ? -2 : -1;
        if (valid && kw == kStrEmpty) {
            // the spare slot holds this one key; its first word is the claim flag (empty -> 0)
            const unsigned long long old = atomicCAS(&table[cap].x, kStrEmpty, 0ull);
            if (old == kStrEmpty) {
                out = publish(&table[cap].y, len, (uint32_t)i);
            } else {
                unsigned long long y;
                do { y = *reinterpret_cast<volatile unsigned long long*>(&table[cap].y); } while (y == kStrPending);
                out = (int64_t)(uint32_t)y;
            }
        } else if (valid) {
            uint64_t s = smix64(kw ^ ((uint64_t)len << 56)) & (cap - 1);
            for (uint64_t probes = 0; probes <= cap; ++probes) {
                // published slots never change, so a cached load is fine: a stale copy can only look emptier
                // than the slot is, and then the CAS (or the volatile re-read) sees the truth
                ulonglong2 sl = *reinterpret_cast<const ulonglong2*>(&table[s]);
                if (sl.x == kStrEmpty) {
                    if (*reinterpret_cast<volatile unsigned long long*>(&counters[1])) break;
                    const unsigned long long old = atomicCAS(&table[s].x, kStrEmpty, kw);
                    if (old == kStrEmpty) {
                        out = publish(&table[s].y, len, (uint32_t)i);
                        break;
                    }
                    sl.x = old;
                    sl.y = kStrPending;
                }
                if (sl.x == kw) {
                    while (sl.y == kStrPending) sl.y = *reinterpret_cast<volatile unsigned long long*>(&table[s].y);
                    if ((uint32_t)(sl.y >> 32) == len) {
                        if (len <= 8) { out = (int64_t)(uint32_t)sl.y; break; }
                        const uint32_t id = (uint32_t)sl.y;
                        const uint32_t r = id <= limit ? *reinterpret_cast<volatile uint32_t*>(&rep[id]) : 0u;
                        if (id <= limit && str_equal(chars, offsets[r], beg, len, total)) { out = (int64_t)id; break; }
                    }
                }
                s = (s + 1) & (cap - 1);
            }
        }
        __syncwarp();
        if (in) st_hint_s64(gid + i, out, stream_pol);      // -2 only after an overflow; the caller retries with a larger table
    }
}

__global__ void fill_slots_kernel(ulonglong2* p, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = make_ulonglong2(kStrEmpty, kStrPending);
}

// representative row of each group id (for the result's key strings)
__global__ void str_group_rows_kernel(const uint32_t* __restrict__ rep, const int64_t* __restrict__ gids,
                                      int64_t ng, int64_t* __restrict__ rows) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ng; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = gids[j];
        rows[j] = g < 0 ? -1 : (int64_t)rep[g];
    }
}

// ---- gather ----
struct RowLenLoad {
    const uint32_t* offsets;
    const int64_t* rows;
    __device__ unsigned long long operator()(int64_t i) const {
        const int64_t r = rows[i];
        return r < 0 ? 0ull : (unsigned long long)(offsets[r + 1] - offsets[r]);
    }
};
struct OffsetStore {
    uint32_t* out;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long) const { out[i] = (uint32_t)excl; }
};

__global__ void __launch_bounds__(kStrThreads) str_copy_kernel(const uint8_t* __restrict__ chars,
                                                               const uint32_t* __restrict__ offsets,
                                                               const int64_t* __restrict__ rows, int64_t m,
                                                               const uint32_t* __restrict__ out_offsets,
                                                               uint8_t* __restrict__ out_chars, uint8_t* __restrict__ out_bitmap,
                                                               const uint8_t* __restrict__ in_bitmap) {
    // one warp per output string: coalesced byte copy
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)kStrThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kStrThreads) >> 5;
    for (int64_t i = warp0; i < m; i += nwarps) {
        const int64_t r = rows[i];
        if (r < 0) continue;
        const uint32_t beg = offsets[r], len = offsets[r + 1] - beg;
        const uint32_t ob = out_offsets[i];
        for (uint32_t b = lane; b < len; b += 32) out_chars[ob + b] = chars[beg + b];
    }
    if (out_bitmap) {
        // bitmap bytes: one thread per output byte
        const int64_t nbytes = (m + 7) >> 3;
        for (int64_t b = blockIdx.x * (int64_t)kStrThreads + threadIdx.x; b < nbytes; b += (int64_t)gridDim.x * kStrThreads) {
            uint8_t v = 0;
            for (int k = 0; k < 8; ++k) {
                const int64_t i = b * 8 + k;
                if (i >= m) break;
                const int64_t r = rows[i];
                const bool valid = r >= 0 && (!in_bitmap || ((in_bitmap[r >> 3] >> (r & 7)) & 1));
                if (valid) v |= (uint8_t)(1u << k);
            }
            out_bitmap[b] = v;
        }
    }
}

// ---- low-cardinality argsort: rank of every distinct string, then a stable radix sort of the ranks ----
__global__ void iota_i64_kernel(int64_t* out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = i;
}
// rank[order[j]] = j: order = the sorted position -> group id permutation of the distinct strings
__global__ void rank_of_order_kernel(const uint64_t* __restrict__ order, int64_t ng, uint32_t* __restrict__ rank) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ng; j += (int64_t)gridDim.x * blockDim.x)
        rank[order[j]] = (uint32_t)j;
}
__global__ void rank_keys_kernel(const int64_t* __restrict__ gid, const uint32_t* __restrict__ rank, int64_t n,
                                 uint32_t* __restrict__ keys) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = rank[gid[i]];
}

}  // namespace

// stable argsort of a device str_arr -> uint32 rows in *rows_out (a scratch-owned buffer pointer is returned
// through the ping-pong arguments, so the caller passes its own buffers)
static int32_t str_argsort_lsd(const uint8_t* chars, const uint32_t* offsets, int64_t n, bool ascending, uint64_t* index_out,
                                  cudaStream_t s) {
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(n < 0xffffffffll, "str_argsort: fewer than 2^32 - 1 rows");
    Scratch meta, keys, kA, kB, pA, pB;
    SDC_TRY(meta.alloc(16, s));
    unsigned init[2] = {0u, 0xffffffffu};
    SDC_CUDA_TRY(cudaMemcpyAsync(meta.p, init, 8, cudaMemcpyHostToDevice, s));
    str_maxlen_kernel<<<grid_for(n, kStrThreads, 8), kStrThreads, 0, s>>>(offsets, n, meta.as<unsigned>());
    SDC_CHECK_LAUNCH();
    unsigned lens[2];
    SDC_CUDA_TRY(cudaMemcpyAsync(lens, meta.p, 8, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    const int nchunks = (int)((lens[0] + 7) / 8);
    SDC_TRY(keys.alloc((size_t)n * 8, s));
    SDC_TRY(kA.alloc((size_t)n * 8, s));
    SDC_TRY(kB.alloc((size_t)n * 8, s));
    SDC_TRY(pA.alloc((size_t)n * 4, s));
    SDC_TRY(pB.alloc((size_t)n * 4, s));
    const uint32_t* perm = nullptr;          // nullptr = identity
    const int grid = grid_for(n, kStrThreads, 8);
    // passes from the least significant key to the most: length, last chunk, ..., chunk 0
    for (int pass = -1; pass < nchunks; ++pass) {
        const int chunk = pass < 0 ? -1 : nchunks - 1 - pass;
        if (chunk < 0 && lens[0] == lens[1]) continue;            // all lengths equal: nothing to order
        str_pass_keys_kernel<<<grid, kStrThreads, 0, s>>>(chars, offsets, perm, n, chunk, keys.as<uint64_t>());
        SDC_CHECK_LAUNCH();
        const void* rk = nullptr;
        const void* rp = nullptr;
        // the pass reads the current perm in its first digit pass only and alternates its outputs
        // starting with the first one given, so the first output must be the buffer perm is NOT in
        const bool in_a = (perm == pA.as<uint32_t>());
        SDC_TRY(radix_sort_core(SDCB200_U64, keys.p, kA.p, kB.p, 4, perm, in_a ? pB.p : pA.p, in_a ? pA.p : pB.p, n,
                                perm == nullptr, !ascending, &rk, &rp, s));
        if (rp == nullptr || rp == (const void*)perm) continue;   // no digit varied: order unchanged
        perm = reinterpret_cast<const uint32_t*>(rp);
    }
    widen_rows_kernel<<<grid, 256, 0, s>>>(perm, index_out, n);
    SDC_CHECK_LAUNCH();
    SDC_CUDA_TRY(cudaStreamSynchronize(s));      // scratch buffers go out of scope
    return SDCB200_OK;
}

}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int32_t sdcb200_str_factorize(const uint8_t* chars, const uint32_t* offsets, const uint8_t* bitmap, int64_t n,
                              int64_t expected_groups, int64_t* gid_out, int64_t* ngroups_out, int64_t* table_handle_words,
                              int64_t* table_cap_out, void* stream);
int32_t sdcb200_str_group_rows(int64_t table_handle_words, int64_t table_cap, const int64_t* gids, int64_t ng, int64_t* rows_out,
                               void* stream);
int32_t sdcb200_str_table_free(int64_t table_handle_words, void* stream);
int32_t sdcb200_str_gather_sizes(const uint32_t* offsets, const int64_t* rows, int64_t m, uint32_t* out_offsets, int64_t* total_chars,
                                 void* stream);
int32_t sdcb200_str_gather_chars(const uint8_t* chars, const uint32_t* offsets, const uint8_t* bitmap, const int64_t* rows, int64_t m,
                                 const uint32_t* out_offsets, uint8_t* out_chars, uint8_t* out_bitmap, void* stream);
}

// Few distinct strings among many rows (C4's 1 K keys over 200 M rows): the LSD sort above moves every row once per
// varying byte (8 passes of 16-byte rows for 8-character keys); instead the rows are factorized (one pass, ids in
// [0, G)), the G distinct strings are sorted by the LSD sort, and the rows are ordered by a stable radix sort of their
// strings' 32-bit RANKS, whose constant high digits cost nothing (2 passes of 8-byte rows for G = 1 K).  Taken when a
// sample of the first rows says every string repeats often enough (see the estimate below); *used = 0 leaves the job
// to the LSD sort.
static int32_t str_argsort_by_rank(const uint8_t* chars, const uint32_t* offsets, int64_t n, bool ascending, uint64_t* index_out,
                                   cudaStream_t s, int32_t* used) {
    *used = 0;
    constexpr int64_t kMinRows = 1 << 20, kSample = 1 << 18;
    if (n < kMinRows || n >= 0xffffffffll) return SDCB200_OK;
    Scratch gid;
    int64_t ng = 0, handle = 0, cap = 0;
    {
        Scratch sample_gid;
        SDC_TRY(sample_gid.alloc((size_t)kSample * 8, s));
        SDC_TRY(sdcb200_str_factorize(chars, offsets, nullptr, kSample, 0, sample_gid.as<int64_t>(), &ng, &handle, &cap, s));
        SDC_TRY(sdcb200_str_table_free(handle, s));
    }
    // d distinct strings among s sampled rows: few (d <= s / 16) means the sample has seen most of them; otherwise the
    // number of distinct strings G of the column is estimated from the collisions inside the sample -- d = G (1 - exp(-s / G))
    // for rows drawn uniformly from G strings, solved by bisection -- and the rank path is taken while the column still
    // repeats every string some 32 times (200 M rows over 1 M strings: factorize + 3 passes over 32-bit ranks against
    // 8 passes over 64-bit chunks).  A sample that is almost collision-free (d > 0.93 s) cannot tell and leaves the job to LSD.
    int64_t expect = ng * 2;
    if (ng * 16 > kSample) {
        const double d = (double)ng, sn = (double)kSample;
        if (d > 0.93 * sn) return SDCB200_OK;
        double lo = d, hi = 64.0 * sn;
        for (int it = 0; it < 60; ++it) {
            const double g = 0.5 * (lo + hi);
            if (g * (1.0 - exp(-sn / g)) < d) lo = g; else hi = g;
        }
        if (hi * 32.0 > (double)n) return SDCB200_OK;
        expect = (int64_t)(hi * 1.25);
    }
    SDC_TRY(gid.alloc((size_t)n * 8, s));
    SDC_TRY(sdcb200_str_factorize(chars, offsets, nullptr, n, expect, gid.as<int64_t>(), &ng, &handle, &cap, s));
    int32_t rc = SDCB200_OK;
    Scratch ids, rows, soff, schars, order, rank, keys, kA, kB, pA, pB;
    auto body = [&]() -> int32_t {
        if (ng * 16 > n) return SDCB200_OK;                          // the sample lied: many distinct strings after all
        // the distinct strings (one representative row each), gathered into a small str_arr and sorted
        SDC_TRY(ids.alloc((size_t)ng * 8, s));
        SDC_TRY(rows.alloc((size_t)ng * 8, s));
        iota_i64_kernel<<<grid_for(ng, 256, 4), 256, 0, s>>>(ids.as<int64_t>(), ng);
        SDC_CHECK_LAUNCH();
        SDC_TRY(sdcb200_str_group_rows(handle, cap, ids.as<int64_t>(), ng, rows.as<int64_t>(), s));
        SDC_TRY(soff.alloc((size_t)(ng + 1) * 4, s));
        int64_t total = 0;
        SDC_TRY(sdcb200_str_gather_sizes(offsets, rows.as<int64_t>(), ng, soff.as<uint32_t>(), &total, s));
        SDC_TRY(schars.alloc((size_t)(total > 0 ? total : 1), s));
        SDC_TRY(sdcb200_str_gather_chars(chars, offsets, nullptr, rows.as<int64_t>(), ng, soff.as<uint32_t>(), schars.as<uint8_t>(),
                                         nullptr, s));
        SDC_TRY(order.alloc((size_t)ng * 8, s));
        SDC_TRY(str_argsort_lsd(schars.as<uint8_t>(), soff.as<uint32_t>(), ng, true, order.as<uint64_t>(), s));
        SDC_TRY(rank.alloc((size_t)ng * 4, s));
        rank_of_order_kernel<<<grid_for(ng, 256, 4), 256, 0, s>>>(order.as<uint64_t>(), ng, rank.as<uint32_t>());
        SDC_CHECK_LAUNCH();
        // rows ordered by the rank of their string: equal strings keep their input order in both directions
        SDC_TRY(keys.alloc((size_t)n * 4, s));
        rank_keys_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(gid.as<int64_t>(), rank.as<uint32_t>(), n, keys.as<uint32_t>());
        SDC_CHECK_LAUNCH();
        SDC_TRY(kA.alloc((size_t)n * 4, s));
        SDC_TRY(kB.alloc((size_t)n * 4, s));
        SDC_TRY(pA.alloc((size_t)n * 4, s));
        SDC_TRY(pB.alloc((size_t)n * 4, s));
        const void *rk = nullptr, *rp = nullptr;
        SDC_TRY(radix_sort_core(SDCB200_U32, keys.p, kA.p, kB.p, 4, nullptr, pA.p, pB.p, n, true, !ascending, &rk, &rp, s));
        widen_rows_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(reinterpret_cast<const uint32_t*>(rp), index_out, n);
        SDC_CHECK_LAUNCH();
        SDC_CUDA_TRY(cudaStreamSynchronize(s));      // scratch buffers go out of scope
        *used = 1;
        return SDCB200_OK;
    };
    rc = body();
    const int32_t rc2 = sdcb200_str_table_free(handle, s);
    return rc != SDCB200_OK ? rc : rc2;
}

extern "C" {

int32_t sdcb200_str_argsort(const uint8_t* chars, const uint32_t* offsets, int64_t n, int32_t ascending, uint64_t* index_out,
                            void* stream) {
    SDC_ARG_CHECK(n >= 0, "n");
    SDC_ARG_CHECK(n == 0 || (offsets && index_out), "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    int32_t used = 0;
    SDC_TRY(str_argsort_by_rank(chars, offsets, n, ascending != 0, index_out, s, &used));
    if (used) return SDCB200_OK;
    return str_argsort_lsd(chars, offsets, n, ascending != 0, index_out, s);
}

int32_t sdcb200_str_factorize(const uint8_t* chars, const uint32_t* offsets, const uint8_t* bitmap, int64_t n,
                              int64_t expected_groups, int64_t* gid_out, int64_t* ngroups_out, int64_t* table_handle_words,
                              int64_t* table_cap_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(n >= 0 && n < 0xffffffffll, "row count (< 2^32 - 1)");
    SDC_ARG_CHECK(ngroups_out && table_handle_words && table_cap_out, "null out pointer");
    *ngroups_out = 0;
    *table_handle_words = 0;
    *table_cap_out = 0;
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(offsets && gid_out, "null pointer");
    uint64_t worst = 64;
    while (worst < 2ull * (uint64_t)n) worst <<= 1;
    // capacity: room for twice the expected groups, eight times while the table stays small (<= 64 K slots = 1 MB)
    uint64_t cap = 1ull << 16;
    if (expected_groups > 0) {
        uint64_t want = 2ull * (uint64_t)expected_groups;
        const uint64_t roomy = 8ull * (uint64_t)expected_groups < (1ull << 16) ? 8ull * (uint64_t)expected_groups : (1ull << 16);
        if (roomy > want) want = roomy;
        cap = 1024;
        while (cap < want) cap <<= 1;
    }
    if (cap > worst) cap = worst;
    for (;;) {
        // one allocation: slots [cap + 1] | counters [2] | rep [limit + 2] (limit = cap / 2 + 1 groups at most)
        const uint64_t limit = cap / 2 + 1;
        void* table = nullptr;
        SDC_TRY(dev_alloc(&table, (size_t)(cap + 1) * 16 + 16 + (size_t)(limit + 2) * 4, s));
        unsigned long long* counters = reinterpret_cast<unsigned long long*>(reinterpret_cast<ulonglong2*>(table) + cap + 1);
        uint32_t* rep = reinterpret_cast<uint32_t*>(counters + 2);
        fill_slots_kernel<<<grid_for((int64_t)cap + 1, 256, 8), 256, 0, s>>>(reinterpret_cast<ulonglong2*>(table), cap + 1);
        SDC_CHECK_LAUNCH();
        SDC_CUDA_TRY(cudaMemsetAsync(counters, 0, 16, s));
        str_factorize_kernel<<<grid_for(n, kStrThreads, 8), kStrThreads, 0, s>>>(chars, offsets, bitmap, n,
                                                                                reinterpret_cast<ulonglong2*>(table), cap, rep,
                                                                                gid_out, counters, limit);
        SDC_CHECK_LAUNCH();
        unsigned long long h[2];
        SDC_CUDA_TRY(cudaMemcpyAsync(h, counters, 16, cudaMemcpyDeviceToHost, s));
        SDC_CUDA_TRY(cudaStreamSynchronize(s));
        if (h[1] == 0) {
            *ngroups_out = (int64_t)h[0];
            *table_handle_words = (int64_t)reinterpret_cast<uintptr_t>(table);
            *table_cap_out = (int64_t)cap;
            return SDCB200_OK;
        }
        dev_free(table, s);
        if (cap >= worst) { set_error("str_factorize: table overflow at capacity %llu", (unsigned long long)cap); return SDCB200_ERR_CAPACITY; }
        cap = cap * 16 < worst ? cap * 16 : worst;
    }
}

int32_t sdcb200_str_group_rows(int64_t table_handle_words, int64_t table_cap, const int64_t* gids, int64_t ng, int64_t* rows_out,
                               void* stream) {
    SDC_ARG_CHECK(ng >= 0 && table_cap > 0, "ng / table capacity");
    if (ng == 0) return SDCB200_OK;
    SDC_ARG_CHECK(table_handle_words && gids && rows_out, "null pointer");
    // rep[] follows the slots and the two counters in the table's allocation (sdcb200_str_factorize)
    const ulonglong2* slots = reinterpret_cast<const ulonglong2*>((uintptr_t)table_handle_words);
    const uint32_t* rep = reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned long long*>(slots + table_cap + 1) + 2);
    str_group_rows_kernel<<<grid_for(ng, 256, 4), 256, 0, (cudaStream_t)stream>>>(rep, gids, ng, rows_out);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

int32_t sdcb200_str_table_free(int64_t table_handle_words, void* stream) {
    if (table_handle_words) return dev_free(reinterpret_cast<void*>((uintptr_t)table_handle_words), (cudaStream_t)stream);
    return SDCB200_OK;
}

int32_t sdcb200_str_gather_sizes(const uint32_t* offsets, const int64_t* rows, int64_t m, uint32_t* out_offsets, int64_t* total_chars,
                                 void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(m >= 0 && total_chars && out_offsets, "args");
    *total_chars = 0;
    if (m == 0) {
        SDC_CUDA_TRY(cudaMemsetAsync(out_offsets, 0, 4, s));
        return SDCB200_OK;
    }
    Scratch tsum;
    SDC_TRY(tsum.alloc((size_t)(ceil_div(m, kScanTile) + 1) * 8, s));
    SDC_TRY(device_exclusive_scan(RowLenLoad{offsets, rows}, OffsetStore{out_offsets}, m, tsum.as<unsigned long long>(), s));
    unsigned long long total = 0;
    SDC_CUDA_TRY(cudaMemcpyAsync(&total, tsum.as<unsigned long long>() + ceil_div(m, kScanTile), 8, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    SDC_ARG_CHECK(total < 0xffffffffull, "gathered chars exceed the uint32 offset range (sdc_str_arr.hpp:206)");
    const uint32_t t32 = (uint32_t)total;
    SDC_CUDA_TRY(cudaMemcpyAsync(out_offsets + m, &t32, 4, cudaMemcpyHostToDevice, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    *total_chars = (int64_t)total;
    return SDCB200_OK;
}

int32_t sdcb200_str_gather_chars(const uint8_t* chars, const uint32_t* offsets, const uint8_t* bitmap, const int64_t* rows, int64_t m,
                                 const uint32_t* out_offsets, uint8_t* out_chars, uint8_t* out_bitmap, void* stream) {
    SDC_ARG_CHECK(m >= 0, "m");
    if (m == 0) return SDCB200_OK;
    str_copy_kernel<<<grid_for(m * 32, kStrThreads, 8), kStrThreads, 0, (cudaStream_t)stream>>>(chars, offsets, rows, m, out_offsets,
                                                                                               out_chars, out_bitmap, bitmap);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

// ------------------------------------------------------------------ string columns in the multi-GPU shuffle
// The reference ships a string column as a LENGTH array plus the characters and rebuilds the offsets on the receiving
// side with a scan (sdc/shuffle_utils.py:277-297, convert_len_arr_to_offset sdc_str_arr.hpp:242-252).  Same wire format
// here; the validity bit rides in bit 31 of the length (a string is shorter than 2 GiB), so no bitmap travels.
//   str_hash          out[i] = 64-bit hash of the string's bytes (NA rows: one fixed value) -- the routing key
//   str_pack_lens     lens[i] = byte length | (NA ? 1 << 31 : 0)
//   str_unpack_lens   lens -> offsets[n + 1] (exclusive scan), validity bitmap, total characters
}  // extern "C"

namespace sdcb200 { namespace {
__global__ void __launch_bounds__(256) str_hash_kernel(const uint8_t* __restrict__ chars, const uint32_t* __restrict__ offsets,
                                                       const uint8_t* __restrict__ bitmap, int64_t n, unsigned long long* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long h = 0xcbf29ce484222325ull;
        if (!bitmap || ((bitmap[i >> 3] >> (i & 7)) & 1)) {
            const uint32_t b = offsets[i], e = offsets[i + 1];
            for (uint32_t j = b; j < e; ++j) h = (h ^ chars[j]) * 0x100000001b3ull;        // FNV-1a over the bytes
            h ^= (unsigned long long)(e - b) << 32;
        } else {
            h = 0x6e756c6c6e756c6cull;
        }
        out[i] = mix64_fin(h);
    }
}
__global__ void __launch_bounds__(256) str_pack_lens_kernel(const uint32_t* __restrict__ offsets, const uint8_t* __restrict__ bitmap,
                                                            int64_t n, uint32_t* __restrict__ lens) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = !bitmap || ((bitmap[i >> 3] >> (i & 7)) & 1);
        lens[i] = (offsets[i + 1] - offsets[i]) | (valid ? 0u : 0x80000000u);
    }
}
struct PackedLenLoad {
    const uint32_t* lens;
    __device__ unsigned long long operator()(int64_t i) const { return lens[i] & 0x7fffffffu; }
};
// one thread per bitmap byte: 8 rows
__global__ void __launch_bounds__(256) str_lens_bitmap_kernel(const uint32_t* __restrict__ lens, int64_t n, uint8_t* __restrict__ bitmap) {
    const int64_t nb = (n + 7) / 8;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
        unsigned v = 0;
        for (int k = 0; k < 8; ++k) {
            const int64_t i = b * 8 + k;
            if (i < n && !(lens[i] >> 31)) v |= 1u << k;
        }
        bitmap[b] = (uint8_t)v;
    }
}

// ---- short string keys as 64-bit words (the hash groupby on a column whose strings all fit one word) ----
// str_profile: shortest / longest string and the number of NA rows -- what decides whether a key column can go through
// the numeric groupby as int64 words instead of being factorized first:
//   * every string exactly 8 bytes, no NA, chars 8-byte aligned: the chars buffer IS the key column (no copy at all;
//     C4's 8-character keys);
//   * every string at most 7 bytes, no NA: word = bytes (little-endian, zero padded) | length << 56 -- the length byte
//     keeps "a" and "a\0" apart, so the word is exactly the string.
__global__ void __launch_bounds__(256) str_profile_kernel(const uint32_t* __restrict__ offsets, const uint8_t* __restrict__ bitmap,
                                                          int64_t n, unsigned long long* __restrict__ out /* [3] min, max, n_na */) {
    unsigned mx = 0, mn = 0xffffffffu;
    unsigned long long na = 0;
    const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
    // 4 rows per step from one aligned 16-byte load (+ the next offset)
    const int64_t n4 = n / 4;
    for (int64_t q = t0; q < n4; q += step) {
        const uint4 o = *reinterpret_cast<const uint4*>(offsets + 4 * q);
        const uint32_t e = offsets[4 * q + 4];
        const unsigned l0 = o.y - o.x, l1 = o.z - o.y, l2 = o.w - o.z, l3 = e - o.w;
        mx = max(max(mx, l0), max(l1, max(l2, l3)));
        mn = min(min(mn, l0), min(l1, min(l2, l3)));
    }
    for (int64_t i = 4 * n4 + t0; i < n; i += step) {
        const unsigned len = offsets[i + 1] - offsets[i];
        mx = max(mx, len);
        mn = min(mn, len);
    }
    if (bitmap) {
        const int64_t nb = (n + 7) / 8;
        for (int64_t b = t0; b < nb; b += step) {
            unsigned v = bitmap[b];
            if (b == nb - 1 && (n & 7)) v |= 0xffu << (n & 7);          // bits beyond row n - 1 do not count
            na += 8u - (unsigned)__popc(v & 0xffu);
        }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    mn = __reduce_min_sync(0xffffffffu, mn);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) na += __shfl_xor_sync(0xffffffffu, na, d);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, (unsigned long long)mn);
        atomicMax(out + 1, (unsigned long long)mx);
        if (na) atomicAdd(out + 2, na);
    }
}

// words[i] = the string's bytes (little-endian, zero padded) | length << 56; lengths are at most 7 (the caller profiled)
__global__ void __launch_bounds__(256) str_pack_words_kernel(const uint8_t* __restrict__ chars, const uint32_t* __restrict__ offsets,
                                                             int64_t n, unsigned long long* __restrict__ words) {
    const uint64_t total = offsets[n];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t beg = offsets[i];
        uint32_t len = offsets[i + 1] - beg;
        len = len > 7u ? 7u : len;
        // the 8 bytes at the aligned address below `beg` and the 8 after them cover any 7 bytes starting at `beg`
        const uint64_t a = (uint64_t)beg & ~7ull;
        const unsigned sh = ((unsigned)beg & 7u) * 8u;
        unsigned long long w = 0;
        if (len) {
            const unsigned long long lo = *reinterpret_cast<const unsigned long long*>(chars + a);
            w = lo >> sh;
            if (sh && a + 8 < total) w |= *reinterpret_cast<const unsigned long long*>(chars + a + 8) << (64u - sh);
            w &= (1ull << (8u * len)) - 1ull;
        }
        words[i] = w | ((unsigned long long)len << 56);
    }
}

// result keys back into a str_arr: lengths (scanned into offsets by the caller's device_exclusive_scan), then the bytes
struct WordLenLoad {
    const unsigned long long* words;
    int fixed8;
    __device__ unsigned long long operator()(int64_t i) const { return fixed8 ? 8ull : (words[i] >> 56); }
};
// order-preserving unsigned image of a key word: byte-lexicographic string order, a proper prefix first
__global__ void __launch_bounds__(256) str_words_sortkeys_kernel(const unsigned long long* __restrict__ words, int64_t m, int fixed8,
                                                                 unsigned long long* __restrict__ keys) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long w = words[i];
        const unsigned long long body = fixed8 ? w : (w & 0x00ffffffffffffffull);
        const unsigned long long be = ((unsigned long long)__byte_perm((unsigned)body, 0u, 0x0123) << 32) |
                                      (unsigned long long)__byte_perm((unsigned)(body >> 32), 0u, 0x0123);
        keys[i] = fixed8 ? be : (be | (w >> 56));        // bytes 0..6 big-endian in the top 56 bits, the length below them
    }
}
__global__ void __launch_bounds__(256) str_fixed8_offsets_kernel(uint32_t* __restrict__ offsets, int64_t m) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= m; i += (int64_t)gridDim.x * blockDim.x)
        offsets[i] = (uint32_t)(8 * i);
}
__global__ void __launch_bounds__(256) str_unpack_words_kernel(const unsigned long long* __restrict__ words, int64_t m, int fixed8,
                                                               const uint32_t* __restrict__ offsets, uint8_t* __restrict__ chars) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long w = words[i];
        const unsigned len = fixed8 ? 8u : (unsigned)(w >> 56);
        const uint32_t ob = offsets[i];
        for (unsigned b = 0; b < len; ++b) chars[ob + b] = (uint8_t)(w >> (8u * b));
    }
}
}}

extern "C" {

int32_t sdcb200_str_profile(const uint32_t* offsets, const uint8_t* bitmap, int64_t n, int64_t* out3, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(n >= 0 && out3 != nullptr, "args");
    out3[0] = out3[1] = out3[2] = 0;
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(offsets != nullptr, "null offsets");
    SDC_ARG_CHECK((reinterpret_cast<uintptr_t>(offsets) & 15) == 0, "offsets must be 16-byte aligned");
    Scratch meta;
    SDC_TRY(meta.alloc(24, s));
    const unsigned long long init[3] = {~0ull, 0ull, 0ull};
    SDC_CUDA_TRY(cudaMemcpyAsync(meta.p, init, 24, cudaMemcpyHostToDevice, s));
    str_profile_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, s>>>(offsets, bitmap, n, meta.as<unsigned long long>());
    SDC_CHECK_LAUNCH();
    unsigned long long h[3];
    SDC_CUDA_TRY(cudaMemcpyAsync(h, meta.p, 24, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    out3[0] = (int64_t)h[0];
    out3[1] = (int64_t)h[1];
    out3[2] = (int64_t)h[2];
    return SDCB200_OK;
}

int32_t sdcb200_str_pack_words(const uint8_t* chars, const uint32_t* offsets, int64_t n, uint64_t* words_out, void* stream) {
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(chars != nullptr && offsets != nullptr && words_out != nullptr, "null pointer");
    SDC_ARG_CHECK((reinterpret_cast<uintptr_t>(chars) & 7) == 0, "chars must be 8-byte aligned");
    str_pack_words_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(chars, offsets, n,
                                                                                reinterpret_cast<unsigned long long*>(words_out));
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

int32_t sdcb200_str_unpack_words(const uint64_t* words, int64_t m, int32_t fixed8, uint32_t* offsets_out, uint8_t* chars_out,
                                 int64_t* total_chars, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(m >= 0 && offsets_out != nullptr && total_chars != nullptr, "args");
    *total_chars = 0;
    if (m == 0) {
        SDC_CUDA_TRY(cudaMemsetAsync(offsets_out, 0, 4, s));
        return SDCB200_OK;
    }
    SDC_ARG_CHECK(words != nullptr && chars_out != nullptr, "null pointer");
    SDC_ARG_CHECK(m < (int64_t)(0xffffffffull / 8), "result chars exceed the uint32 offset range (sdc_str_arr.hpp:206)");
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(words);
    if (fixed8) {
        // every string is 8 bytes: offsets and total are known without a scan or a read-back
        str_fixed8_offsets_kernel<<<grid_for(m + 1, 256, 8), 256, 0, s>>>(offsets_out, m);
        SDC_CHECK_LAUNCH();
        SDC_CUDA_TRY(cudaMemcpyAsync(chars_out, w, (size_t)m * 8, cudaMemcpyDeviceToDevice, s));
        *total_chars = 8 * m;
        return SDCB200_OK;
    }
    Scratch tsum;
    SDC_TRY(tsum.alloc((size_t)(ceil_div(m, kScanTile) + 1) * 8, s));
    SDC_TRY(device_exclusive_scan(WordLenLoad{w, fixed8}, OffsetStore{offsets_out}, m, tsum.as<unsigned long long>(), s));
    str_unpack_words_kernel<<<grid_for(m, 256, 8), 256, 0, s>>>(w, m, fixed8, offsets_out, chars_out);
    SDC_CHECK_LAUNCH();
    unsigned long long total = 0;
    SDC_CUDA_TRY(cudaMemcpyAsync(&total, tsum.as<unsigned long long>() + ceil_div(m, kScanTile), 8, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    const uint32_t t32 = (uint32_t)total;
    SDC_CUDA_TRY(cudaMemcpyAsync(offsets_out + m, &t32, 4, cudaMemcpyHostToDevice, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    *total_chars = (int64_t)total;
    return SDCB200_OK;
}

int32_t sdcb200_str_words_order(const uint64_t* words, int64_t m, int32_t fixed8, int64_t* order_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(m >= 0, "m");
    if (m == 0) return SDCB200_OK;
    SDC_ARG_CHECK(words != nullptr && order_out != nullptr, "null pointer");
    Scratch keys;
    SDC_TRY(keys.alloc((size_t)m * 8, s));
    str_words_sortkeys_kernel<<<grid_for(m, 256, 8), 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(words), m, fixed8,
                                                                keys.as<unsigned long long>());
    SDC_CHECK_LAUNCH();
    SDC_TRY(argsort_device(SDCB200_U64, keys.p, m, true, order_out, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));      // the key scratch goes out of scope
    return SDCB200_OK;
}

int32_t sdcb200_str_hash(const uint8_t* chars, const uint32_t* offsets, const uint8_t* bitmap, int64_t n, uint64_t* out, void* stream) {
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(offsets != nullptr && out != nullptr, "null pointer");
    str_hash_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(chars, offsets, bitmap, n, reinterpret_cast<unsigned long long*>(out));
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

int32_t sdcb200_str_pack_lens(const uint32_t* offsets, const uint8_t* bitmap, int64_t n, uint32_t* lens_out, void* stream) {
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(offsets != nullptr && lens_out != nullptr, "null pointer");
    str_pack_lens_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(offsets, bitmap, n, lens_out);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

int32_t sdcb200_str_unpack_lens(const uint32_t* lens, int64_t n, uint32_t* offsets_out, uint8_t* bitmap_out, int64_t* total_chars,
                                void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(n >= 0 && offsets_out != nullptr && total_chars != nullptr, "args");
    *total_chars = 0;
    if (n == 0) {
        SDC_CUDA_TRY(cudaMemsetAsync(offsets_out, 0, 4, s));
        return SDCB200_OK;
    }
    SDC_ARG_CHECK(lens != nullptr, "null lens");
    Scratch tsum;
    SDC_TRY(tsum.alloc((size_t)(ceil_div(n, kScanTile) + 1) * 8, s));
    SDC_TRY(device_exclusive_scan(PackedLenLoad{lens}, OffsetStore{offsets_out}, n, tsum.as<unsigned long long>(), s));
    if (bitmap_out) {
        str_lens_bitmap_kernel<<<grid_for((n + 7) / 8, 256, 1), 256, 0, s>>>(lens, n, bitmap_out);
        SDC_CHECK_LAUNCH();
    }
    unsigned long long total = 0;
    SDC_CUDA_TRY(cudaMemcpyAsync(&total, tsum.as<unsigned long long>() + ceil_div(n, kScanTile), 8, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    SDC_ARG_CHECK(total < 0xffffffffull, "received chars exceed the uint32 offset range (sdc_str_arr.hpp:206)");
    const uint32_t t32 = (uint32_t)total;
    SDC_CUDA_TRY(cudaMemcpyAsync(offsets_out + n, &t32, 4, cudaMemcpyHostToDevice, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    *total_chars = (int64_t)total;
    return SDCB200_OK;
}

// The reference's own symbol (sdc/native/str_ext/sdc_str_arr.hpp:565-593, bound in sdc/str_arr_ext.py:94-99):
// HOST buffers, void return.
void stable_argsort(const char* data, const uint32_t* offsets, uint64_t len, int8_t ascending, uint64_t* result) {
    if (len == 0) return;
    uint32_t total = offsets[len];
    void *dchars = nullptr, *doffs = nullptr, *dout = nullptr;
    cudaStream_t s = nullptr;
    auto fail = [&](const char* what) {
        fprintf(stderr, "sdc_b200 stable_argsort: %s: %s\n", what, sdcb200_last_error());
    };
    if (dev_alloc(&dchars, (size_t)total + 16, s) || dev_alloc(&doffs, (size_t)(len + 1) * 4, s) ||
        dev_alloc(&dout, (size_t)len * 8, s)) { fail("alloc"); return; }
    bool ok = cudaMemcpyAsync(dchars, data, total, cudaMemcpyHostToDevice, s) == cudaSuccess &&
              cudaMemcpyAsync(doffs, offsets, (size_t)(len + 1) * 4, cudaMemcpyHostToDevice, s) == cudaSuccess;
    if (ok) ok = sdcb200_str_argsort((const uint8_t*)dchars, (const uint32_t*)doffs, (int64_t)len, ascending, (uint64_t*)dout, s) == SDCB200_OK;
    if (ok) ok = cudaMemcpyAsync(result, dout, (size_t)len * 8, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
                 cudaStreamSynchronize(s) == cudaSuccess;
    if (!ok) fail("run");
    dev_free(dchars, s); dev_free(doffs, s); dev_free(dout, s);
}

}  // extern "C"
