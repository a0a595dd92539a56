// join.cu -- pd.merge(how='inner' | 'left') hash join on one key column (K7 in SURVEY.md).
//
// The reference snapshot has no merge overload (sdc/hiframes/join.py:34-43 is a stub and every test
// in sdc/tests/test_join.py:51-394 is skipped), so the behaviour follows pandas.merge, and the
// artefact follows the reference's live indexer convention: two int64 row-position arrays with -1
// for "no partner" and the full cross product for duplicate keys
// (_sdc_internal_join, sdc/datatypes/common_functions.py:225-455); payload columns are gathered by
// the caller (sdcb200_gather / sdcb200_take_nullable_f64 = setitem_arr_nan, hiframes/join.py:34-43).
//
//   build  (right side)  open-addressing table of 16-byte slots {key, val} in HBM, linear probing,
//                        64-bit atomicCAS on the key word; val counts the rows of the key
//                        (the atomicAdd's return value is the row's rank inside its key group);
//                        an exclusive scan over the slots turns val into (start << 32 | count) and a
//                        scatter lays the right row ids out grouped by key: rid[start + rank] = row.
//   probe  (left side)   pass A: one 16-byte slot load per probe step gives start and count together;
//                        per-row val is kept (8 B) and per-tile output counts are reduced;
//                        a scan of the tile counts gives every tile its output offset (and n_out);
//                        pass B: block scan of the per-row counts, then (left row, right row) pairs are
//                        written in LEFT ROW ORDER (how='left' keeps the left order like pandas).
// Keys: int64, or float64 canonicalised so that -0.0 == +0.0 and NaN == NaN (pandas merges NaN keys).
// Limits: build side < 2^32 - 1 rows.
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "scan.cuh"

namespace sdcb200 {
namespace {

constexpr unsigned long long kEmptyKey = 0x7ff8dead0000beefull;   // never a canonical float key; an int64 key
                                                                  // equal to it lives in the spare slot [cap]
constexpr int kJoinThreads = kScanThreads;         // 256
constexpr int kItems = kScanItems;                 // 8 rows per thread in the probe tiles
constexpr int kTile = kScanTile;                   // 2048 rows per CTA

__host__ __device__ inline unsigned long long jmix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <bool KEY_F64>
__device__ __forceinline__ unsigned long long canon_key(unsigned long long b) {
    if (KEY_F64) {
        const unsigned long long mag = b & 0x7fffffffffffffffull;
        if (mag > 0x7ff0000000000000ull) return 0x7ff8000000000000ull;   // one NaN
        if (mag == 0) return 0ull;                                       // -0.0 == +0.0
    }
    return b;
}

struct JTable {
    ulonglong2* slots;          // [cap + 1]
    unsigned long long cap;     // power of two
};

__device__ __forceinline__ unsigned long long jt_upsert(const JTable& t, unsigned long long key) {
    if (key == kEmptyKey) return t.cap;
    unsigned long long s = jmix64(key) & (t.cap - 1);
    while (true) {
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&t.slots[s].x);
        if (cur == key) return s;
        if (cur == kEmptyKey) {
            const unsigned long long old = atomicCAS(&t.slots[s].x, kEmptyKey, key);
            if (old == kEmptyKey || old == key) return s;
        }
        s = (s + 1) & (t.cap - 1);
    }
}

// -> val of the key's slot, 0 when the key is absent
__device__ __forceinline__ unsigned long long jt_lookup(const JTable& t, unsigned long long key) {
    if (key == kEmptyKey) return t.slots[t.cap].y;
    unsigned long long s = jmix64(key) & (t.cap - 1);
    while (true) {
        const ulonglong2 sl = __ldg(reinterpret_cast<const ulonglong2*>(&t.slots[s]));
        if (sl.x == key) return sl.y;
        if (sl.x == kEmptyKey) return 0ull;
        s = (s + 1) & (t.cap - 1);
    }
}

__global__ void jt_init_kernel(JTable t) {
    const unsigned long long n = t.cap + 1;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        t.slots[i] = make_ulonglong2(kEmptyKey, 0ull);
}

template <bool KEY_F64>
__global__ void __launch_bounds__(kJoinThreads) jt_insert_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                                                 JTable t, uint32_t* __restrict__ rslot,
                                                                 uint32_t* __restrict__ rrank) {
    const int64_t stride = (int64_t)gridDim.x * kJoinThreads;
    for (int64_t j = (int64_t)blockIdx.x * kJoinThreads + threadIdx.x; j < n; j += stride) {
        const unsigned long long k = canon_key<KEY_F64>(keys[j]);
        const unsigned long long s = jt_upsert(t, k);
        const unsigned long long rank = atomicAdd(&t.slots[s].y, 1ull);
        rslot[j] = (uint32_t)s;
        rrank[j] = (uint32_t)rank;
    }
}

struct SlotCountLoad {
    const ulonglong2* slots;
    __device__ unsigned long long operator()(int64_t i) const { return slots[i].y; }
};
struct SlotStartStore {          // val = start << 32 | count
    ulonglong2* slots;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        slots[i].y = (excl << 32) | v;
    }
};
__global__ void __launch_bounds__(kJoinThreads) jt_scatter_kernel(const uint32_t* __restrict__ rslot,
                                                                  const uint32_t* __restrict__ rrank, int64_t n, JTable t,
                                                                  uint32_t* __restrict__ rid) {
    const int64_t stride = (int64_t)gridDim.x * kJoinThreads;
    for (int64_t j = (int64_t)blockIdx.x * kJoinThreads + threadIdx.x; j < n; j += stride) {
        const unsigned long long val = t.slots[rslot[j]].y;
        rid[(val >> 32) + rrank[j]] = (uint32_t)j;
    }
}

// ---- probe pass A: per-row val, per-tile output count ----
template <bool KEY_F64, bool LEFT>
__global__ void __launch_bounds__(kJoinThreads) probe_count_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                                                   JTable t, unsigned long long* __restrict__ pval,
                                                                   unsigned long long* __restrict__ tile_counts) {
    __shared__ unsigned long long wsum[kJoinThreads / 32];
    const int64_t i0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
    unsigned long long k[kItems], v[kItems];
    if (i0 + kItems <= n) {
#pragma unroll
        for (int e = 0; e < kItems; e += 2) {
            const ulonglong2 kk = __ldg(reinterpret_cast<const ulonglong2*>(keys + i0 + e));
            k[e] = kk.x;
            k[e + 1] = kk.y;
        }
    } else {
#pragma unroll
        for (int e = 0; e < kItems; ++e) k[e] = (i0 + e < n) ? keys[i0 + e] : 0ull;
    }
    unsigned long long c = 0;
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        v[e] = (i0 + e < n) ? jt_lookup(t, canon_key<KEY_F64>(k[e])) : 0ull;
        const unsigned long long cnt = v[e] & 0xffffffffull;
        if (i0 + e < n) c += LEFT ? (cnt ? cnt : 1ull) : cnt;
    }
    if (i0 + kItems <= n) {
#pragma unroll
        for (int e = 0; e < kItems; e += 2)
            *reinterpret_cast<ulonglong2*>(pval + i0 + e) = make_ulonglong2(v[e], v[e + 1]);
    } else {
#pragma unroll
        for (int e = 0; e < kItems; ++e)
            if (i0 + e < n) pval[i0 + e] = v[e];
    }
    unsigned long long tot;
    block_excl_scan_u64(c, &tot, wsum);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}

// ---- probe pass B: emit (left row, right row) pairs in left-row order ----
template <bool LEFT>
__global__ void __launch_bounds__(kJoinThreads) probe_emit_kernel(const unsigned long long* __restrict__ pval, int64_t n,
                                                                  const unsigned long long* __restrict__ tile_offs,
                                                                  const uint32_t* __restrict__ rid,
                                                                  int64_t* __restrict__ lidx, int64_t* __restrict__ ridx) {
    __shared__ unsigned long long wsum[kJoinThreads / 32];
    const int64_t i0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
    unsigned long long v[kItems];
    if (i0 + kItems <= n) {
#pragma unroll
        for (int e = 0; e < kItems; e += 2) {
            const ulonglong2 vv = __ldg(reinterpret_cast<const ulonglong2*>(pval + i0 + e));
            v[e] = vv.x;
            v[e + 1] = vv.y;
        }
    } else {
#pragma unroll
        for (int e = 0; e < kItems; ++e) v[e] = (i0 + e < n) ? pval[i0 + e] : 0ull;
    }
    unsigned long long c = 0;
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        const unsigned long long cnt = v[e] & 0xffffffffull;
        if (i0 + e < n) c += LEFT ? (cnt ? cnt : 1ull) : cnt;
    }
    unsigned long long tot;
    unsigned long long pos = tile_offs[blockIdx.x] + block_excl_scan_u64(c, &tot, wsum);
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        if (i0 + e >= n) break;
        const unsigned long long cnt = v[e] & 0xffffffffull, start = v[e] >> 32;
        if (cnt == 1) {
            lidx[pos] = i0 + e;
            ridx[pos] = (int64_t)rid[start];
            ++pos;
        } else if (cnt == 0) {
            if (LEFT) {
                lidx[pos] = i0 + e;
                ridx[pos] = -1;
                ++pos;
            }
        } else {
            for (unsigned long long m = 0; m < cnt; ++m) {
                lidx[pos + m] = i0 + e;
                ridx[pos + m] = (int64_t)rid[start + m];
            }
            pos += cnt;
        }
    }
}

// dst[i] = idx[i] < 0 ? NaN : (double)src[idx[i]]  -- the payload gather of a left join
// (setitem_arr_nan, sdc/hiframes/join.py:34-43: unmatched rows become NaN, ints are promoted)
template <typename T>
__global__ void take_nullable_kernel(const T* __restrict__ src, const int64_t* __restrict__ idx, double* __restrict__ dst,
                                     int64_t n) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = __ldg(idx + i);
        dst[i] = j < 0 ? nan : (double)src[j];
    }
}

struct JoinResult {
    int64_t n_out = 0;
    int64_t* lidx = nullptr;
    int64_t* ridx = nullptr;
};

std::mutex g_jmu;
std::unordered_map<int64_t, JoinResult*> g_joins;
int64_t g_next_join = 1;

template <bool KEY_F64, bool LEFT>
int32_t run_join(const unsigned long long* l, int64_t nl, const unsigned long long* r, int64_t nr, cudaStream_t s,
                 JoinResult* res) {
    const int sms = sm_count();
    // ---- build ----
    JTable t;
    t.cap = 1024;
    while (t.cap * 3 < (unsigned long long)nr * 4) t.cap <<= 1;        // load factor <= 0.75
    Scratch slots, rslot, rrank, rid, tsum_b;
    SDC_TRY(slots.alloc((size_t)(t.cap + 1) * 16, s));
    t.slots = slots.as<ulonglong2>();
    SDC_TRY(rslot.alloc((size_t)nr * 4, s));
    SDC_TRY(rrank.alloc((size_t)nr * 4, s));
    SDC_TRY(rid.alloc((size_t)nr * 4, s));
    const int64_t nslots = (int64_t)t.cap + 1;
    SDC_TRY(tsum_b.alloc((size_t)(ceil_div(nslots, kTile) + 1) * 8, s));
    const int wide = sms * 8;
    jt_init_kernel<<<(unsigned)(ceil_div(nslots, 256) < wide ? ceil_div(nslots, 256) : wide), 256, 0, s>>>(t);
    SDC_CHECK_LAUNCH();
    if (nr > 0) {
        const int grid = (int)(ceil_div(nr, kJoinThreads) < wide ? ceil_div(nr, kJoinThreads) : wide);
        jt_insert_kernel<KEY_F64><<<grid, kJoinThreads, 0, s>>>(r, nr, t, rslot.as<uint32_t>(), rrank.as<uint32_t>());
        SDC_CHECK_LAUNCH();
        SDC_TRY(device_exclusive_scan(SlotCountLoad{t.slots}, SlotStartStore{t.slots}, nslots,
                                      tsum_b.as<unsigned long long>(), s));
        jt_scatter_kernel<<<grid, kJoinThreads, 0, s>>>(rslot.as<uint32_t>(), rrank.as<uint32_t>(), nr, t, rid.as<uint32_t>());
        SDC_CHECK_LAUNCH();
    }
    // ---- probe ----
    if (nl == 0) return SDCB200_OK;
    const int64_t nt = ceil_div(nl, kTile);
    Scratch pval, tcnt;
    SDC_TRY(pval.alloc((size_t)nl * 8, s));
    SDC_TRY(tcnt.alloc((size_t)(nt + 1) * 8 + (size_t)(ceil_div(nt, kTile) + 1) * 8, s));
    unsigned long long* tile_counts = tcnt.as<unsigned long long>();
    unsigned long long* tsum_p = tile_counts + nt + 1;
    probe_count_kernel<KEY_F64, LEFT><<<(unsigned)nt, kJoinThreads, 0, s>>>(l, nl, t, pval.as<unsigned long long>(), tile_counts);
    SDC_CHECK_LAUNCH();
    // exclusive scan of the tile counts in place; total lands in tile_counts[nt]
    if (nt <= 64 * 1024) {
        scan_tiles_kernel<<<1, kJoinThreads, 0, s>>>(tile_counts, nt);
        SDC_CHECK_LAUNCH();
    } else {
        SDC_TRY(device_exclusive_scan(ArrayLoad{tile_counts}, ArrayStore{tile_counts}, nt, tsum_p, s));
        // the total of the tile sums is tsum_p[ceil(nt / kTile)]
        SDC_CUDA_TRY(cudaMemcpyAsync(tile_counts + nt, tsum_p + ceil_div(nt, kTile), 8, cudaMemcpyDeviceToDevice, s));
    }
    unsigned long long total = 0;
    SDC_CUDA_TRY(cudaMemcpyAsync(&total, tile_counts + nt, 8, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    res->n_out = (int64_t)total;
    if (total == 0) return SDCB200_OK;
    void* p = nullptr;
    SDC_TRY(dev_alloc(&p, (size_t)total * 8, s));
    res->lidx = reinterpret_cast<int64_t*>(p);
    SDC_TRY(dev_alloc(&p, (size_t)total * 8, s));
    res->ridx = reinterpret_cast<int64_t*>(p);
    probe_emit_kernel<LEFT><<<(unsigned)nt, kJoinThreads, 0, s>>>(pval.as<unsigned long long>(), nl, tile_counts,
                                                                  rid.as<uint32_t>(), res->lidx, res->ridx);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

void free_join(JoinResult* r, cudaStream_t s) {
    if (!r) return;
    if (r->lidx) dev_free(r->lidx, s);
    if (r->ridx) dev_free(r->ridx, s);
    delete r;
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int32_t sdcb200_join(int32_t how, int32_t key_dtype, const void* left_keys, int64_t nl, const void* right_keys, int64_t nr,
                     void* stream, int64_t* handle_out, int64_t* n_out) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(how == SDCB200_JOIN_INNER || how == SDCB200_JOIN_LEFT, "how must be INNER or LEFT");
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(nl >= 0 && nr >= 0 && nr < 0xffffffffll, "row counts (build side < 2^32 - 1 rows)");
    SDC_ARG_CHECK((nl == 0 || left_keys) && (nr == 0 || right_keys), "null key column");
    SDC_ARG_CHECK(handle_out && n_out, "null out pointer");
    JoinResult* res = new JoinResult();
    const unsigned long long* l = reinterpret_cast<const unsigned long long*>(left_keys);
    const unsigned long long* r = reinterpret_cast<const unsigned long long*>(right_keys);
    int32_t rc;
    const bool f = key_dtype == SDCB200_F64, left = how == SDCB200_JOIN_LEFT;
    if (f) rc = left ? run_join<true, true>(l, nl, r, nr, s, res) : run_join<true, false>(l, nl, r, nr, s, res);
    else rc = left ? run_join<false, true>(l, nl, r, nr, s, res) : run_join<false, false>(l, nl, r, nr, s, res);
    if (rc != SDCB200_OK) {
        free_join(res, s);
        return rc;
    }
    std::lock_guard<std::mutex> lk(g_jmu);
    const int64_t h = g_next_join++;
    g_joins[h] = res;
    *handle_out = h;
    *n_out = res->n_out;
    return SDCB200_OK;
}

int32_t sdcb200_join_indexers(int64_t handle, int64_t* left_idx, int64_t* right_idx, int32_t dst_is_host, void* stream) {
    JoinResult* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_jmu);
        auto it = g_joins.find(handle);
        if (it != g_joins.end()) r = it->second;
    }
    if (!r) { set_error("join: unknown handle %lld", (long long)handle); return SDCB200_ERR_HANDLE; }
    if (r->n_out == 0) return SDCB200_OK;
    const cudaMemcpyKind kind = dst_is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    cudaStream_t s = (cudaStream_t)stream;
    if (left_idx) SDC_CUDA_TRY(cudaMemcpyAsync(left_idx, r->lidx, (size_t)r->n_out * 8, kind, s));
    if (right_idx) SDC_CUDA_TRY(cudaMemcpyAsync(right_idx, r->ridx, (size_t)r->n_out * 8, kind, s));
    if (dst_is_host) SDC_CUDA_TRY(cudaStreamSynchronize(s));
    return SDCB200_OK;
}

int32_t sdcb200_join_free(int64_t handle, void* stream) {
    JoinResult* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_jmu);
        auto it = g_joins.find(handle);
        if (it == g_joins.end()) { set_error("join: unknown handle %lld", (long long)handle); return SDCB200_ERR_HANDLE; }
        r = it->second;
        g_joins.erase(it);
    }
    free_join(r, (cudaStream_t)stream);
    return SDCB200_OK;
}

int32_t sdcb200_take_nullable_f64(int32_t src_dtype, const void* src, const int64_t* idx, int64_t n, double* dst,
                                  void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(src_dtype == SDCB200_I64 || src_dtype == SDCB200_F64, "source dtype must be I64 or F64");
    SDC_ARG_CHECK(n >= 0, "n");
    if (n == 0) return SDCB200_OK;
    const int grid = (int)(ceil_div(n, 256) < (int64_t)sm_count() * 16 ? ceil_div(n, 256) : (int64_t)sm_count() * 16);
    if (src_dtype == SDCB200_F64)
        take_nullable_kernel<double><<<grid, 256, 0, s>>>(reinterpret_cast<const double*>(src), idx, dst, n);
    else
        take_nullable_kernel<long long><<<grid, 256, 0, s>>>(reinterpret_cast<const long long*>(src), idx, dst, n);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

}  // extern "C"
