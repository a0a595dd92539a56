// groupby.cu -- hash groupby-aggregate (K5, K8 in SURVEY.md).
//
// Replaces the reference's dict-of-row-lists groupby
//   build  sdc/datatypes/hpat_pandas_dataframe_functions.py:3076-3097
//   merge  sdc/datatypes/hpat_pandas_groupby_functions.py:58-73
//   per-group take + Series.<agg>()  groupby_functions.py:194-243
// with a one-pass hash aggregate: results are the same (keys exact, count/min/max exact,
// sum/mean/var/std to rounding), the algorithm is not.
//
//   * keys: int64 or float64 (NaN keys skipped, -0.0 == +0.0, reference :3088-3089).
//   * open-addressing table, linear probing, 64-bit atomicCAS on the key word.
//   * LOW CARDINALITY (fits a shared-memory table): every CTA aggregates its rows into a
//     private shared-memory table (native u32 / CAS-loop f64 shared atomics), then flushes the
//     occupied slots into the global table -- HBM traffic = keys + values read once.
//     Rows whose key does not fit the shared table fall through to the global table, so the
//     path is correct for any cardinality and merely slower.
//   * HIGH CARDINALITY: rows go straight to the global table (L2 / HBM atomics, RED.ADD.F64).
//   * aggregates kept per (column, slot): sum, count of non-NaN, min, max, first row (sort=False);
//     var/std run a second pass over the rows for sum((x-mean)^2) like nanvar's two passes
//     (sdc/functions/numpy_like.py:850-872).
//   * ordering (K8): occupied slots are compacted, group keys argsorted with the library's own
//     radix sort (sort=True, groupby_functions.py:209-210) or by first row (sort=False = the
//     dict's first-appearance order, :58-73), then every aggregate is gathered in that order.
//
// Skipna semantics: sum of an all-NaN group = 0, mean = NaN, min/max = NaN,
// count = 0; +-inf take part in min/max/sum; integer columns: sum/min/max/count are int64.
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "radix_sort.cuh"

namespace sdcb200 {
namespace {

// empty-slot marker: a NaN payload, never a canonical float key; an int64 key equal to it uses the spare slot
__host__ __device__ constexpr unsigned long long kEmptyKey() { return 0x7ff8dead0000beefull; }
constexpr int kMaxCols = 32;
constexpr int kGbThreads = 256;

enum : unsigned { AGG_SUM = 1, AGG_MEAN = 2, AGG_MIN = 4, AGG_MAX = 8, AGG_COUNT = 16, AGG_VAR = 32, AGG_STD = 64 };
enum { NEED_SUM = 1, NEED_CNT = 2, NEED_MIN = 4, NEED_MAX = 8, NEED_FIRST = 16 };

__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// order-preserving map double <-> u64 for atomicMin/atomicMax
__device__ __forceinline__ unsigned long long f64_to_ordered(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_to_f64(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned long long i64_to_ordered(long long v) {
    return (unsigned long long)v ^ 0x8000000000000000ull;
}

// The global table.  `nrep` replicas of (cap + 1) slots each: the shared-memory path flushes CTA b
// into replica b % nrep so that the flush atomics of hundreds of CTAs do not serialise on the same
// few addresses; replicas 1.. are folded into replica 0 afterwards.  Slot indices below are
// "global": replica * (cap + 1) + slot.
struct Table {
    unsigned long long* keys;       // [nrep][cap + 1]; slot cap = the spare slot for key == kEmpty
    unsigned long long* first;      // first row of the group (NEED_FIRST)
    // per column c (nullptr when not needed)
    void* sum[kMaxCols];            // double or long long
    unsigned long long* cnt[kMaxCols];
    unsigned long long* mn[kMaxCols];   // ordered u64
    unsigned long long* mx[kMaxCols];
    double* ssd[kMaxCols];          // second pass: sum of squared deviations (replica 0 only)
    unsigned long long cap;         // power of two
    unsigned long long limit;       // max occupied slots per replica before "overflow"
    unsigned long long* occupied;   // [nrep] device counters
    int* overflow;                  // device flag
    int nrep;
};

struct Cols {
    const void* data[kMaxCols];
    int is_f64[kMaxCols];
    int ncols;
    int vec_ok;                     // key and value columns are 16-byte aligned: 128-bit loads
};

template <bool KEY_F64>
__device__ __forceinline__ bool canon_key(unsigned long long b, unsigned long long& k) {
    if (KEY_F64) {
        const unsigned long long mag = b & 0x7fffffffffffffffull;
        if (mag > 0x7ff0000000000000ull) return false;   // NaN key: row skipped (reference :3088-3089)
        if (mag == 0) b = 0;                             // -0.0 == +0.0
    }
    k = b;
    return true;
}

// find-or-insert in replica `rep`; returns the global slot, or ~0 when the table overflowed
__device__ __forceinline__ unsigned long long global_upsert(const Table& t, int rep, unsigned long long key) {
    const unsigned long long off = (unsigned long long)rep * (t.cap + 1);
    if (key == kEmptyKey()) {
        if (atomicCAS(&t.keys[off + t.cap], kEmptyKey(), key + 1) == kEmptyKey()) atomicAdd(&t.occupied[rep], 1ull);
        return off + t.cap;
    }
    unsigned long long slot = mix64(key) & (t.cap - 1);
    for (unsigned long long probes = 0; probes <= t.cap; ++probes) {
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&t.keys[off + slot]);
        if (cur == key) return off + slot;
        if (cur == kEmptyKey()) {
            const unsigned long long old = atomicCAS(&t.keys[off + slot], kEmptyKey(), key);
            if (old == kEmptyKey()) {
                if (atomicAdd(&t.occupied[rep], 1ull) + 1 > t.limit) { *(volatile int*)t.overflow = 1; return ~0ull; }
                return off + slot;
            }
            if (old == key) return off + slot;
        }
        slot = (slot + 1) & (t.cap - 1);
    }
    *(volatile int*)t.overflow = 1;
    return ~0ull;
}

// read-only lookup in replica 0 (second pass); the key is known to be present
__device__ __forceinline__ unsigned long long global_find(const Table& t, unsigned long long key) {
    if (key == kEmptyKey()) return t.cap;
    unsigned long long slot = mix64(key) & (t.cap - 1);
    while (t.keys[slot] != key) slot = (slot + 1) & (t.cap - 1);
    return slot;
}

// one value (raw bits) of column c into global slot g
__device__ __forceinline__ void global_accumulate(const Table& t, int c, bool is_f64, int need, unsigned long long g,
                                                  unsigned long long bits) {
    if (is_f64) {
        const double v = __longlong_as_double((long long)bits);
        if (v != v) return;                                    // skipna
        if (need & NEED_SUM) atomicAdd(reinterpret_cast<double*>(t.sum[c]) + g, v);
        if (need & NEED_CNT) atomicAdd(&t.cnt[c][g], 1ull);
        if (need & NEED_MIN) atomicMin(&t.mn[c][g], f64_to_ordered(v));
        if (need & NEED_MAX) atomicMax(&t.mx[c][g], f64_to_ordered(v));
    } else {
        if (need & NEED_SUM) atomicAdd(reinterpret_cast<unsigned long long*>(t.sum[c]) + g, bits);
        if (need & NEED_CNT) atomicAdd(&t.cnt[c][g], 1ull);
        if (need & NEED_MIN) atomicMin(&t.mn[c][g], i64_to_ordered((long long)bits));
        if (need & NEED_MAX) atomicMax(&t.mx[c][g], i64_to_ordered((long long)bits));
    }
}

// ---- tile loads: a CTA works on kGbTile = 256 threads x 4 rows; thread t holds rows
// base + 2t, 2t+1, 512 + 2t, 512 + 2t + 1, so each column costs two coalesced 128-bit loads per
// thread and all loads of a tile are in flight together.
constexpr int kRows = 4;
constexpr int kGbTile = kGbThreads * kRows;

__device__ __forceinline__ int64_t tile_row(int64_t base, int tid, int e) {
    return base + (e >> 1) * (2 * kGbThreads) + 2 * tid + (e & 1);
}

__device__ __forceinline__ void load4(const void* col, int64_t base, int64_t n, int tid, bool vec,
                                      unsigned long long v[kRows]) {
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(col);
    if (vec && base + kGbTile <= n) {
        const int4 a = ld_stream_v4(p + base + 2 * tid);
        const int4 b = ld_stream_v4(p + base + 2 * kGbThreads + 2 * tid);
        v[0] = (unsigned long long)(unsigned)a.x | ((unsigned long long)(unsigned)a.y << 32);
        v[1] = (unsigned long long)(unsigned)a.z | ((unsigned long long)(unsigned)a.w << 32);
        v[2] = (unsigned long long)(unsigned)b.x | ((unsigned long long)(unsigned)b.y << 32);
        v[3] = (unsigned long long)(unsigned)b.z | ((unsigned long long)(unsigned)b.w << 32);
    } else {
#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            const int64_t r = tile_row(base, tid, e);
            v[e] = r < n ? p[r] : 0ull;
        }
    }
}

// ------------------------------------------------------------------ high-cardinality path
template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_global_kernel(const void* __restrict__ keys, int64_t n, Cols cols,
                                                               Table t, int need) {
    const int tid = threadIdx.x;
    const int64_t ntiles = (n + kGbTile - 1) / kGbTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * kGbTile;
        unsigned long long kb[kRows], v0[kRows];
        load4(keys, base, n, tid, cols.vec_ok, kb);
        if (cols.ncols > 0) load4(cols.data[0], base, n, tid, cols.vec_ok, v0);
        if (*(volatile int*)t.overflow) return;
        unsigned long long g[kRows];
#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            unsigned long long k;
            g[e] = ~0ull;
            if (tile_row(base, tid, e) < n && canon_key<KEY_F64>(kb[e], k)) g[e] = global_upsert(t, 0, k);
        }
        if (need & NEED_FIRST) {
#pragma unroll
            for (int e = 0; e < kRows; ++e)
                if (g[e] != ~0ull) atomicMin(&t.first[g[e]], (unsigned long long)tile_row(base, tid, e));
        }
        for (int c = 0; c < cols.ncols; ++c) {
            if (c > 0) load4(cols.data[c], base, n, tid, cols.vec_ok, v0);
            const bool f = cols.is_f64[c] != 0;
#pragma unroll
            for (int e = 0; e < kRows; ++e)
                if (g[e] != ~0ull) global_accumulate(t, c, f, need, g[e], v0[e]);
        }
    }
}

// ------------------------------------------------------------------ low-cardinality path
// Per-CTA shared-memory table of SCAP slots; layout (SoA):
//   keys[SCAP] u64 | first[SCAP] u64 | per column: sum[SCAP] 8B, mn[SCAP] u64, mx[SCAP] u64 | cnt[SCAP] u32 per column
// (only the arrays `need` asks for are carved out).  Rows that cannot get a slot within
// kSmemProbes probes go to the global table directly.  HBM traffic = keys + values read once.
constexpr int kSmemProbes = 16;

struct SmemLayout {
    int scap;
    int off_first;          // byte offsets from the table base; -1 = absent
    int off_sum[kMaxCols], off_cnt[kMaxCols], off_mn[kMaxCols], off_mx[kMaxCols];
    int bytes;
};

// strided sample of the key column -> dense-key window for the direct-addressed mode of gb_smem_kernel:
// direct[0] = lo, direct[1] = range (0 = use hashing).  int64 keys only.
constexpr int kSampleThreads = 1024;
constexpr int kSamplePerThread = 4;
__global__ void __launch_bounds__(kSampleThreads) gb_sample_kernel(const long long* __restrict__ keys, int64_t n, int scap,
                                                                   long long* direct) {
    __shared__ long long smin[kSampleThreads / 32], smax[kSampleThreads / 32];
    const int64_t total = (int64_t)kSampleThreads * kSamplePerThread;
    const int64_t stride = n > total ? n / total : 1;
    long long mn = 0x7fffffffffffffffLL, mx = -0x7fffffffffffffffLL - 1;
    long long kv[kSamplePerThread];
#pragma unroll
    for (int e = 0; e < kSamplePerThread; ++e) {        // independent loads, all in flight together
        const int64_t r = ((int64_t)e * kSampleThreads + threadIdx.x) * stride;
        kv[e] = r < n ? __ldg(keys + r) : keys[0];
    }
#pragma unroll
    for (int e = 0; e < kSamplePerThread; ++e) {
        mn = kv[e] < mn ? kv[e] : mn;
        mx = kv[e] > mx ? kv[e] : mx;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, mn, d), b = __shfl_xor_sync(0xffffffffu, mx, d);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kSampleThreads / 32; ++w) {
            mn = smin[w] < mn ? smin[w] : mn;
            mx = smax[w] > mx ? smax[w] : mx;
        }
        // work in the order-preserving unsigned image of int64 so that extreme keys cannot overflow
        const unsigned long long top = 0x8000000000000000ull;
        const unsigned long long umn = (unsigned long long)mn ^ top, umx = (unsigned long long)mx ^ top;
        long long lo = 0, range = 0;
        if (mx >= mn && umx - umn < (unsigned long long)scap) {
            // centre the scap-wide window on the sampled keys, clamped to the int64 range
            unsigned long long shift = ((unsigned long long)scap - (umx - umn) - 1) / 2;
            if (shift > umn) shift = umn;
            unsigned long long ulo = umn - shift;
            if (ulo > ~0ull - (unsigned long long)scap) ulo = ~0ull - (unsigned long long)scap;
            lo = (long long)(ulo ^ top);
            range = scap;
        }
        direct[0] = lo;
        direct[1] = range;
    }
}

template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_smem_kernel(const void* __restrict__ keys, int64_t n, Cols cols, Table t,
                                                             int need, SmemLayout lay, const long long* __restrict__ direct) {
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sm);
    unsigned long long* sfirst = reinterpret_cast<unsigned long long*>(sm + (lay.off_first >= 0 ? lay.off_first : 0));
    const int scap = lay.scap;
    const int tid = threadIdx.x;
    const int rep = blockIdx.x % t.nrep;
    // dense int64 keys: slot = key - lo, no hashing, no probing (window chosen by gb_sample_kernel)
    const unsigned long long dlo = KEY_F64 ? 0ull : (unsigned long long)direct[0];
    const unsigned long long drange = KEY_F64 ? 0ull : (unsigned long long)direct[1];
    for (int i = tid; i < scap; i += kGbThreads) {
        skeys[i] = kEmptyKey();
        if (need & NEED_FIRST) sfirst[i] = ~0ull;
        for (int c = 0; c < cols.ncols; ++c) {
            if (need & NEED_SUM) reinterpret_cast<unsigned long long*>(sm + lay.off_sum[c])[i] = 0ull;
            if (need & NEED_CNT) reinterpret_cast<unsigned*>(sm + lay.off_cnt[c])[i] = 0u;
            if (need & NEED_MIN) reinterpret_cast<unsigned long long*>(sm + lay.off_mn[c])[i] = ~0ull;
            if (need & NEED_MAX) reinterpret_cast<unsigned long long*>(sm + lay.off_mx[c])[i] = 0ull;
        }
    }
    __syncthreads();
    const int64_t ntiles = (n + kGbTile - 1) / kGbTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncwarp();
        const int64_t base = tile * kGbTile;
        unsigned long long kb[kRows], v0[kRows];
        load4(keys, base, n, tid, cols.vec_ok, kb);
        if (cols.ncols > 0) load4(cols.data[0], base, n, tid, cols.vec_ok, v0);
        if (*(volatile int*)t.overflow) break;
        int slot[kRows];                      // >= 0 shared slot, -1 row skipped, -2 row lives in the global table
        unsigned long long g[kRows];
#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            unsigned long long k = 0;
            const bool valid = tile_row(base, tid, e) < n && canon_key<KEY_F64>(kb[e], k);
            int sl = -1;
            g[e] = ~0ull;
            bool pending = valid && k != kEmptyKey();
            if (drange) {
                const unsigned long long d = k - dlo;
                if (pending && d < drange) {
                    sl = (int)d;
                    if (skeys[sl] != k) skeys[sl] = k;          // presence marker; every writer stores the same value
                }
            } else {
                unsigned s = (unsigned)((k * 0x9E3779B97F4A7C15ull) >> 40) & (unsigned)(scap - 1);
                // the warp probes in lock step: the loop exit is warp-uniform, no divergence to undo
                for (int p = 0; p < kSmemProbes; ++p) {
                    if (!__any_sync(0xffffffffu, pending)) break;
                    if (pending) {
                        const unsigned long long cur = skeys[s];
                        if (cur == k) { sl = (int)s; pending = false; }
                        else if (cur == kEmptyKey()) {
                            const unsigned long long old = atomicCAS(&skeys[s], kEmptyKey(), k);
                            if (old == kEmptyKey() || old == k) { sl = (int)s; pending = false; }
                            else s = (s + 1) & (unsigned)(scap - 1);
                        } else s = (s + 1) & (unsigned)(scap - 1);
                    }
                }
            }
            if (valid && sl < 0) {      // outside the window / shared table full around this key: global table
                g[e] = global_upsert(t, rep, k);
                sl = g[e] != ~0ull ? -2 : -1;
            }
            slot[e] = sl;
            __syncwarp();
        }
        if (need & NEED_FIRST) {
#pragma unroll
            for (int e = 0; e < kRows; ++e) {
                const unsigned long long row = (unsigned long long)tile_row(base, tid, e);
                if (slot[e] >= 0) { if (row < sfirst[slot[e]]) atomicMin(&sfirst[slot[e]], row); }
                else if (slot[e] == -2) atomicMin(&t.first[g[e]], row);
            }
            __syncwarp();
        }
        for (int c = 0; c < cols.ncols; ++c) {
            if (c > 0) load4(cols.data[c], base, n, tid, cols.vec_ok, v0);
            const bool f = cols.is_f64[c] != 0;
            double* ssum = reinterpret_cast<double*>(sm + lay.off_sum[c]);
            unsigned* scnt = reinterpret_cast<unsigned*>(sm + lay.off_cnt[c]);
            unsigned long long* smn = reinterpret_cast<unsigned long long*>(sm + lay.off_mn[c]);
            unsigned long long* smx = reinterpret_cast<unsigned long long*>(sm + lay.off_mx[c]);
#pragma unroll
            for (int e = 0; e < kRows; ++e) {
                const int sl = slot[e];
                if (sl == -2) global_accumulate(t, c, f, need, g[e], v0[e]);
                if (sl >= 0) {
                    if (f) {
                        const double v = __longlong_as_double((long long)v0[e]);
                        if (v == v) {
                            if (need & NEED_SUM) atomicAdd(ssum + sl, v);
                            if (need & NEED_CNT) atomicAdd(scnt + sl, 1u);
                            if (need & NEED_MIN) atomicMin(smn + sl, f64_to_ordered(v));
                            if (need & NEED_MAX) atomicMax(smx + sl, f64_to_ordered(v));
                        }
                    } else {
                        if (need & NEED_SUM) atomicAdd(reinterpret_cast<unsigned long long*>(ssum) + sl, v0[e]);
                        if (need & NEED_CNT) atomicAdd(scnt + sl, 1u);
                        if (need & NEED_MIN) atomicMin(smn + sl, i64_to_ordered((long long)v0[e]));
                        if (need & NEED_MAX) atomicMax(smx + sl, i64_to_ordered((long long)v0[e]));
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    // flush the occupied shared slots into this CTA's replica of the global table
    for (int i = tid; i < scap; i += kGbThreads) {
        const unsigned long long k = skeys[i];
        if (k == kEmptyKey()) continue;
        const unsigned long long g = global_upsert(t, rep, k);
        if (g == ~0ull) continue;
        if (need & NEED_FIRST) atomicMin(&t.first[g], sfirst[i]);
        for (int c = 0; c < cols.ncols; ++c) {
            if (need & NEED_CNT) {
                const unsigned cn = reinterpret_cast<unsigned*>(sm + lay.off_cnt[c])[i];
                if (cn) atomicAdd(&t.cnt[c][g], (unsigned long long)cn);
            }
            if (need & NEED_SUM) {
                if (cols.is_f64[c]) {
                    const double v = reinterpret_cast<double*>(sm + lay.off_sum[c])[i];
                    if (v != 0.0) atomicAdd(reinterpret_cast<double*>(t.sum[c]) + g, v);
                } else {
                    atomicAdd(reinterpret_cast<unsigned long long*>(t.sum[c]) + g,
                              reinterpret_cast<unsigned long long*>(sm + lay.off_sum[c])[i]);
                }
            }
            if (need & NEED_MIN) atomicMin(&t.mn[c][g], reinterpret_cast<unsigned long long*>(sm + lay.off_mn[c])[i]);
            if (need & NEED_MAX) atomicMax(&t.mx[c][g], reinterpret_cast<unsigned long long*>(sm + lay.off_mx[c])[i]);
        }
    }
}

// fold replicas 1 .. nrep-1 into replica 0
__global__ void gb_merge_replicas_kernel(Table t, int ncols, unsigned is_f64_mask, int need) {
    const unsigned long long per = t.cap + 1;
    const unsigned long long total = per * (unsigned long long)(t.nrep - 1);
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long src = per + i;
        const unsigned long long kraw = t.keys[src];
        if (kraw == kEmptyKey()) continue;
        const bool spare = (src % per) == t.cap;
        const unsigned long long g = global_upsert(t, 0, spare ? kEmptyKey() : kraw);
        if (g == ~0ull) continue;
        if (need & NEED_FIRST) atomicMin(&t.first[g], t.first[src]);
        for (int c = 0; c < ncols; ++c) {
            if (need & NEED_CNT) { const unsigned long long cn = t.cnt[c][src]; if (cn) atomicAdd(&t.cnt[c][g], cn); }
            if (need & NEED_SUM) {
                if ((is_f64_mask >> c) & 1u) {
                    const double v = reinterpret_cast<const double*>(t.sum[c])[src];
                    if (v != 0.0) atomicAdd(reinterpret_cast<double*>(t.sum[c]) + g, v);
                } else {
                    atomicAdd(reinterpret_cast<unsigned long long*>(t.sum[c]) + g,
                              reinterpret_cast<const unsigned long long*>(t.sum[c])[src]);
                }
            }
            if (need & NEED_MIN) atomicMin(&t.mn[c][g], t.mn[c][src]);
            if (need & NEED_MAX) atomicMax(&t.mx[c][g], t.mx[c][src]);
        }
    }
}

// second pass for var / std: ssd[slot] += (x - mean)^2  (nanvar's two passes, numpy_like.py:850-872)
template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_ssd_kernel(const void* __restrict__ keys, int64_t n, Cols cols, Table t) {
    const int tid = threadIdx.x;
    const int64_t ntiles = (n + kGbTile - 1) / kGbTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * kGbTile;
        unsigned long long kb[kRows], v0[kRows], g[kRows];
        load4(keys, base, n, tid, cols.vec_ok, kb);
#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            unsigned long long k;
            g[e] = ~0ull;
            if (tile_row(base, tid, e) < n && canon_key<KEY_F64>(kb[e], k)) g[e] = global_find(t, k);
        }
        for (int c = 0; c < cols.ncols; ++c) {
            load4(cols.data[c], base, n, tid, cols.vec_ok, v0);
#pragma unroll
            for (int e = 0; e < kRows; ++e) {
                if (g[e] == ~0ull) continue;
                double v, mean;
                const double cnt = (double)t.cnt[c][g[e]];
                if (cols.is_f64[c]) {
                    v = __longlong_as_double((long long)v0[e]);
                    if (v != v) continue;
                    mean = reinterpret_cast<const double*>(t.sum[c])[g[e]] / cnt;
                } else {
                    v = (double)(long long)v0[e];
                    mean = (double)reinterpret_cast<const long long*>(t.sum[c])[g[e]] / cnt;
                }
                const double d = v - mean;
                atomicAdd(&t.ssd[c][g[e]], d * d);
            }
        }
    }
}

// every table array gets its initial value in one launch
struct InitPlan {
    unsigned long long* ptr[4 + 5 * kMaxCols];
    unsigned long long val[4 + 5 * kMaxCols];
    int count;
};
__global__ void gb_init_kernel(InitPlan plan, unsigned long long n) {
    for (int a = blockIdx.y; a < plan.count; a += gridDim.y) {
        unsigned long long* p = plan.ptr[a];
        const unsigned long long v = plan.val[a];
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
             i += (unsigned long long)gridDim.x * blockDim.x)
            p[i] = v;
    }
}

// ---- result values --------------------------------------------------------------------------
// Output layout: keys[ngcap] and vals[(rank(agg) * ncols + c) * ngcap + j], rank = position of the
// aggregate among the requested ones (ascending bit order).
__device__ __forceinline__ unsigned long long agg_value(const Table& t, bool f, int c, unsigned agg, long long ddof,
                                                        unsigned long long slot) {
    const unsigned long long nanb = 0x7ff8000000000000ull;
    const unsigned long long cn = t.cnt[c] ? t.cnt[c][slot] : 0ull;
    switch (agg) {
        case AGG_SUM:
            return reinterpret_cast<const unsigned long long*>(t.sum[c])[slot];
        case AGG_MEAN: {
            const double s = f ? reinterpret_cast<const double*>(t.sum[c])[slot]
                               : (double)reinterpret_cast<const long long*>(t.sum[c])[slot];
            return cn ? (unsigned long long)__double_as_longlong(s / (double)cn) : nanb;
        }
        case AGG_MIN:
        case AGG_MAX: {
            const unsigned long long o = agg == AGG_MIN ? t.mn[c][slot] : t.mx[c][slot];
            if (f) return cn ? (unsigned long long)__double_as_longlong(ordered_to_f64(o)) : nanb;
            return o ^ 0x8000000000000000ull;
        }
        case AGG_COUNT:
            return cn;
        case AGG_VAR:
        case AGG_STD: {
            if ((long long)cn <= ddof) return nanb;
            const double dn = (double)cn;
            double v = (t.ssd[c][slot] / dn) * dn / (double)((long long)cn - ddof);
            if (agg == AGG_STD) v = sqrt(v);
            return (unsigned long long)__double_as_longlong(v);
        }
    }
    return nanb;
}

__device__ __forceinline__ void write_group(const Table& t, int ncols, unsigned is_f64_mask, unsigned agg_mask,
                                            long long ddof, unsigned long long slot, int64_t j, int64_t ngcap,
                                            unsigned long long* out_keys, unsigned long long* out_vals) {
    out_keys[j] = slot == t.cap ? kEmptyKey() : t.keys[slot];
    int rank = 0;
    for (unsigned bit = 1; bit <= AGG_STD; bit <<= 1) {
        if (!(agg_mask & bit)) continue;
        for (int c = 0; c < ncols; ++c)
            out_vals[((int64_t)rank * ncols + c) * ngcap + j] = agg_value(t, (is_f64_mask >> c) & 1u, c, bit, ddof, slot);
        ++rank;
    }
}

// ---- small tables (cap + 1 <= kSmallSlots): ONE CTA compacts the occupied slots, orders them with a
// bitonic sort in shared memory and writes every aggregate -- no host round trip, no radix sort.
constexpr int kSmallSlots = 8192 + 1;
constexpr int kSmallSort = 8192;          // groups <= limit = cap / 2 + 1 <= 4097 fit twice over
constexpr int kSmallThreads = 1024;

__global__ void __launch_bounds__(kSmallThreads) gb_finalize_small_kernel(Table t, int ncols, unsigned is_f64_mask,
                                                                          unsigned agg_mask, long long ddof, int by_first,
                                                                          int key_f64, int64_t ngcap,
                                                                          unsigned long long* out_keys,
                                                                          unsigned long long* out_vals,
                                                                          unsigned long long* ng_out) {
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(sm);          // [kSmallSort]
    unsigned* sslot = reinterpret_cast<unsigned*>(skey + kSmallSort);              // [kSmallSort]
    __shared__ unsigned counter;
    if (threadIdx.x == 0) counter = 0;
    __syncthreads();
    const unsigned nslots = (unsigned)t.cap + 1;
    constexpr int kPer = (kSmallSlots + kSmallThreads - 1) / kSmallThreads;
    unsigned long long kk[kPer];
#pragma unroll
    for (int e = 0; e < kPer; ++e) {               // independent loads, all in flight together
        const unsigned i = e * kSmallThreads + threadIdx.x;
        kk[e] = i < nslots ? t.keys[i] : kEmptyKey();
    }
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        const unsigned i = e * kSmallThreads + threadIdx.x;
        const unsigned long long k = kk[e];
        if (k == kEmptyKey()) continue;
        const unsigned j = atomicAdd(&counter, 1u);
        if (j >= (unsigned)kSmallSort) continue;
        unsigned long long sk;
        if (by_first) sk = t.first[i];
        else {
            const unsigned long long key = (i == (unsigned)t.cap) ? kEmptyKey() : k;
            sk = key_f64 ? f64_to_ordered(__longlong_as_double((long long)key)) : i64_to_ordered((long long)key);
        }
        skey[j] = sk;
        sslot[j] = i;
    }
    __syncthreads();
    const unsigned ng = counter < (unsigned)kSmallSort ? counter : (unsigned)kSmallSort;
    unsigned P = 1;
    while (P < ng) P <<= 1;
    for (unsigned i = ng + threadIdx.x; i < P; i += kSmallThreads) { skey[i] = ~0ull; sslot[i] = 0xffffffffu; }
    __syncthreads();
    for (unsigned k = 2; k <= P; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < P; i += kSmallThreads) {
                const unsigned ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = skey[i], b = skey[ixj];
                    const unsigned sa = sslot[i], sb = sslot[ixj];
                    const bool gt = a > b || (a == b && sa > sb);
                    const bool up = (i & k) == 0;
                    if (gt == up) { skey[i] = b; skey[ixj] = a; sslot[i] = sb; sslot[ixj] = sa; }
                }
            }
            __syncthreads();
        }
    }
    for (unsigned j = threadIdx.x; j < ng; j += kSmallThreads)
        write_group(t, ncols, is_f64_mask, agg_mask, ddof, sslot[j], j, ngcap, out_keys, out_vals);
    if (threadIdx.x == 0) *ng_out = ng;
}

// ---- large tables: compaction, radix argsort of the group keys, gather ----
__global__ void gb_compact_kernel(Table t, unsigned long long* slots, unsigned long long* sortkeys, int by_first,
                                  unsigned long long* counter) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i <= t.cap;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = t.keys[i];
        if (k == kEmptyKey()) continue;
        const unsigned long long j = atomicAdd(counter, 1ull);
        slots[j] = i;
        sortkeys[j] = by_first ? t.first[i] : (i == t.cap ? kEmptyKey() : k);
    }
}

__global__ void gb_finalize_kernel(Table t, int ncols, unsigned is_f64_mask, unsigned agg_mask, long long ddof,
                                   const unsigned long long* __restrict__ slots, const int64_t* __restrict__ order,
                                   int64_t ng, unsigned long long* out_keys, unsigned long long* out_vals) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ng; j += (int64_t)gridDim.x * blockDim.x)
        write_group(t, ncols, is_f64_mask, agg_mask, ddof, slots[order[j]], j, ng, out_keys, out_vals);
}

// ------------------------------------------------------------------ host side
struct GroupbyResult {
    int key_dtype = 0;
    int ncols = 0;
    int64_t ngroups = 0;
    int64_t ngcap = 0;                 // row stride of `vals`
    unsigned agg_mask = 0;
    unsigned long long* keys = nullptr;    // [ngcap]
    unsigned long long* vals = nullptr;    // [popc(agg_mask) * ncols * ngcap]
};

std::mutex g_mu;
std::unordered_map<int64_t, GroupbyResult*> g_results;
int64_t g_next_handle = 1;

void free_result(GroupbyResult* r) {
    if (!r) return;
    if (r->keys) dev_free(r->keys, nullptr);
    if (r->vals) dev_free(r->vals, nullptr);
    delete r;
}

int need_for(unsigned agg_mask, bool sort) {
    int need = 0;
    if (agg_mask & (AGG_SUM | AGG_MEAN | AGG_VAR | AGG_STD)) need |= NEED_SUM;
    if (agg_mask & (AGG_MEAN | AGG_MIN | AGG_MAX | AGG_COUNT | AGG_VAR | AGG_STD)) need |= NEED_CNT;
    if (agg_mask & AGG_MIN) need |= NEED_MIN;
    if (agg_mask & AGG_MAX) need |= NEED_MAX;
    if (!sort) need |= NEED_FIRST;
    return need;
}

SmemLayout make_layout(int ncols, int need, int want_slots, int budget_bytes) {
    SmemLayout L;
    int per_slot = 8 + ((need & NEED_FIRST) ? 8 : 0);
    per_slot += ncols * (((need & NEED_SUM) ? 8 : 0) + ((need & NEED_CNT) ? 4 : 0) + ((need & NEED_MIN) ? 8 : 0) +
                         ((need & NEED_MAX) ? 8 : 0));
    int scap = 256;
    while (scap < want_slots && scap < 4096) scap <<= 1;
    while (scap > 64 && scap * per_slot > budget_bytes) scap >>= 1;
    L.scap = scap;
    int off = scap * 8;
    L.off_first = -1;
    if (need & NEED_FIRST) { L.off_first = off; off += scap * 8; }
    for (int c = 0; c < kMaxCols; ++c) L.off_sum[c] = L.off_cnt[c] = L.off_mn[c] = L.off_mx[c] = 0;
    for (int c = 0; c < ncols; ++c) {
        if (need & NEED_SUM) { L.off_sum[c] = off; off += scap * 8; }
        if (need & NEED_MIN) { L.off_mn[c] = off; off += scap * 8; }
        if (need & NEED_MAX) { L.off_mx[c] = off; off += scap * 8; }
    }
    for (int c = 0; c < ncols; ++c)
        if (need & NEED_CNT) { L.off_cnt[c] = off; off += scap * 4; }
    L.bytes = off;
    return L;
}

// one stream-ordered allocation per attempt, carved into the table arrays (256-byte aligned pieces)
struct TableMem {
    Scratch block;
    char* cur = nullptr;
    char* end = nullptr;
    int32_t reserve(size_t bytes, cudaStream_t s) {
        SDC_TRY(block.alloc(bytes, s));
        cur = block.as<char>();
        end = cur + bytes;
        return SDCB200_OK;
    }
    int32_t get(void** p, size_t bytes, cudaStream_t) {
        const size_t b = (bytes + 255) & ~size_t(255);
        if (cur + b > end) { set_error("groupby: table arena exhausted"); return SDCB200_ERR_NOMEM; }
        *p = cur;
        cur += b;
        return SDCB200_OK;
    }
};

struct Attempt {
    bool smem;
    unsigned long long cap;
    int nrep;
};

unsigned long long pow2_at_least(unsigned long long v, unsigned long long lo) {
    unsigned long long c = lo;
    while (c < v) c <<= 1;
    return c;
}

template <bool KEY_F64>
int32_t run_groupby(const void* keys, int64_t n, Cols cols, unsigned agg_mask, bool sort, long long ddof,
                    int key_dtype, int64_t expected_groups, cudaStream_t s, GroupbyResult* res) {
    const int need = need_for(agg_mask, sort);
    const bool need_ssd = (agg_mask & (AGG_VAR | AGG_STD)) != 0;
    const int sms = sm_count();
    const int nagg = __builtin_popcount(agg_mask);
    unsigned is_f64_mask = 0;
    for (int c = 0; c < cols.ncols; ++c) if (cols.is_f64[c]) is_f64_mask |= 1u << c;
    cols.vec_ok = (reinterpret_cast<uintptr_t>(keys) & 15) == 0;
    for (int c = 0; c < cols.ncols; ++c) if (reinterpret_cast<uintptr_t>(cols.data[c]) & 15) cols.vec_ok = 0;

    // plan: shared-memory pre-aggregation + small replicated table when the cardinality is (or may be)
    // low; a single global table sized from the hint; the worst case (every row its own group) last
    const unsigned long long worst = pow2_at_least(2ull * (unsigned long long)n, 64);
    Attempt plan[3];
    int nattempts = 0;
    if (expected_groups <= 0) {
        if (n >= 4096) plan[nattempts++] = {true, 8192, 8};
        if ((1ull << 22) < worst) plan[nattempts++] = {false, 1ull << 22, 1};
    } else if (expected_groups <= 4096 && n >= 4096) {
        plan[nattempts++] = {true, pow2_at_least(2ull * (unsigned long long)expected_groups, 1024), 8};
    } else {
        const unsigned long long c = pow2_at_least(2ull * (unsigned long long)expected_groups, 1024);
        if (c < worst) plan[nattempts++] = {false, c, 1};
    }
    plan[nattempts++] = {false, worst, 1};

    for (int attempt = 0; attempt < nattempts; ++attempt) {
        const Attempt& at = plan[attempt];
        TableMem mem;
        Table t;
        memset(&t, 0, sizeof(t));
        t.cap = at.cap;
        t.limit = t.cap / 2 + 1;
        t.nrep = at.nrep;
        const unsigned long long per = t.cap + 1;
        const unsigned long long slots_n = per * (unsigned long long)t.nrep;
        InitPlan ip;
        ip.count = 0;
        void* p;
        {
            int narr = 1 + ((need & NEED_FIRST) ? 1 : 0);
            narr += cols.ncols * (((need & NEED_SUM) ? 1 : 0) + ((need & NEED_CNT) ? 1 : 0) + ((need & NEED_MIN) ? 1 : 0) +
                                  ((need & NEED_MAX) ? 1 : 0) + (need_ssd ? 1 : 0));
            SDC_TRY(mem.reserve((size_t)narr * ((slots_n * 8 + 255) & ~size_t(255)) + 256, s));
        }
        auto add = [&](void** dst, unsigned long long init) -> int32_t {
            SDC_TRY(mem.get(&p, slots_n * 8, s));
            *dst = p;
            ip.ptr[ip.count] = reinterpret_cast<unsigned long long*>(p);
            ip.val[ip.count++] = init;
            return SDCB200_OK;
        };
        SDC_TRY(add(reinterpret_cast<void**>(&t.keys), kEmptyKey()));
        if (need & NEED_FIRST) SDC_TRY(add(reinterpret_cast<void**>(&t.first), ~0ull));
        for (int c = 0; c < cols.ncols; ++c) {
            if (need & NEED_SUM) SDC_TRY(add(&t.sum[c], 0ull));
            if (need & NEED_CNT) SDC_TRY(add(reinterpret_cast<void**>(&t.cnt[c]), 0ull));
            if (need & NEED_MIN) SDC_TRY(add(reinterpret_cast<void**>(&t.mn[c]), ~0ull));
            if (need & NEED_MAX) SDC_TRY(add(reinterpret_cast<void**>(&t.mx[c]), 0ull));
            if (need_ssd) SDC_TRY(add(reinterpret_cast<void**>(&t.ssd[c]), 0ull));
        }
        SDC_TRY(mem.get(&p, 256, s));
        SDC_CUDA_TRY(cudaMemsetAsync(p, 0, 256, s));
        t.occupied = reinterpret_cast<unsigned long long*>(p);                // [nrep <= 8]
        t.overflow = reinterpret_cast<int*>(reinterpret_cast<char*>(p) + 64);
        unsigned long long* dmeta = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(p) + 128);   // ng / compact counter
        {
            unsigned gx = (unsigned)(ceil_div((int64_t)slots_n, 1024) < (int64_t)sms * 4 ? ceil_div((int64_t)slots_n, 1024) : (int64_t)sms * 4);
            if (gx == 0) gx = 1;
            dim3 grid(gx, ip.count < 16 ? ip.count : 16);
            gb_init_kernel<<<grid, 256, 0, s>>>(ip, slots_n);
            SDC_CHECK_LAUNCH();
        }
        const int64_t ntiles = ceil_div(n, kGbTile);
        if (at.smem) {
            SmemLayout lay = make_layout(cols.ncols, need, (int)(2 * (expected_groups > 0 ? expected_groups : 2048)), 160 * 1024);
            auto kern = gb_smem_kernel<KEY_F64>;
            static thread_local bool configured = false;
            if (!configured) {
                SDC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                configured = true;
            }
            static thread_local int occ_bytes = -1, occ_cached = 1;
            if (occ_bytes != lay.bytes) {
                SDC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cached, kern, kGbThreads, lay.bytes));
                occ_bytes = lay.bytes;
            }
            int occ = occ_cached;
            if (occ < 1) occ = 1;
            if (occ > 6) occ = 6;
            int64_t grid = (int64_t)sms * occ;
            if (grid > ntiles) grid = ntiles;
            long long* direct = reinterpret_cast<long long*>(dmeta + 2);      // zeroed: range 0 = hashing
            if (!KEY_F64) {
                gb_sample_kernel<<<1, kSampleThreads, 0, s>>>(reinterpret_cast<const long long*>(keys), n, lay.scap, direct);
                SDC_CHECK_LAUNCH();
            }
            kern<<<(unsigned)grid, kGbThreads, lay.bytes, s>>>(keys, n, cols, t, need, lay, direct);
            SDC_CHECK_LAUNCH();
            if (t.nrep > 1) {
                const int64_t total = (int64_t)per * (t.nrep - 1);
                gb_merge_replicas_kernel<<<(unsigned)(ceil_div(total, 256) < (int64_t)sms * 4 ? ceil_div(total, 256) : (int64_t)sms * 4), 256, 0, s>>>(
                    t, cols.ncols, is_f64_mask, need);
                SDC_CHECK_LAUNCH();
            }
        } else {
            int64_t grid = (int64_t)sms * 8;
            if (grid > ntiles) grid = ntiles;
            gb_global_kernel<KEY_F64><<<(unsigned)grid, kGbThreads, 0, s>>>(keys, n, cols, t, need);
            SDC_CHECK_LAUNCH();
        }
        const bool small = per <= (unsigned long long)kSmallSlots;
        if (small && !need_ssd) {
            // everything else happens in one CTA; the group count comes back with the overflow flag
            const int64_t ngcap = (int64_t)t.limit + 1;
            SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->keys), (size_t)ngcap * 8, s));
            SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->vals), (size_t)ngcap * 8 * (nagg * cols.ncols + 1), s));
            res->ngcap = ngcap;
            auto fk = gb_finalize_small_kernel;
            static thread_local bool fconfigured = false;
            const int fsmem = kSmallSort * 12;
            if (!fconfigured) {
                SDC_CUDA_TRY(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, fsmem));
                fconfigured = true;
            }
            fk<<<1, kSmallThreads, fsmem, s>>>(t, cols.ncols, is_f64_mask, agg_mask, ddof, sort ? 0 : 1, KEY_F64 ? 1 : 0,
                                               ngcap, res->keys, res->vals, dmeta);
            SDC_CHECK_LAUNCH();
            unsigned long long hmeta[17];
            SDC_CUDA_TRY(cudaMemcpyAsync(hmeta, t.occupied, 136, cudaMemcpyDeviceToHost, s));
            SDC_CUDA_TRY(cudaStreamSynchronize(s));
            if ((int)(hmeta[8] & 0xffffffffull)) {
                dev_free(res->keys, s); dev_free(res->vals, s);
                res->keys = res->vals = nullptr;
                if (attempt + 1 < nattempts) continue;
                set_error("groupby: hash table overflow at capacity %llu", t.cap);
                return SDCB200_ERR_CAPACITY;
            }
            res->ngroups = (int64_t)hmeta[16];
            return SDCB200_OK;
        }
        // large table (or var / std, which need a second pass over the rows once the means are known)
        unsigned long long hmeta[9];
        SDC_CUDA_TRY(cudaMemcpyAsync(hmeta, t.occupied, 72, cudaMemcpyDeviceToHost, s));
        SDC_CUDA_TRY(cudaStreamSynchronize(s));
        if ((int)(hmeta[8] & 0xffffffffull)) {
            if (attempt + 1 < nattempts) continue;
            set_error("groupby: hash table overflow at capacity %llu", t.cap);
            return SDCB200_ERR_CAPACITY;
        }
        if (need_ssd) {
            int64_t grid = (int64_t)sms * 8;
            if (grid > ntiles) grid = ntiles;
            gb_ssd_kernel<KEY_F64><<<(unsigned)grid, kGbThreads, 0, s>>>(keys, n, cols, t);
            SDC_CHECK_LAUNCH();
        }
        const int64_t ng = (int64_t)hmeta[0];        // replica 0 holds every group after the merge
        res->ngroups = ng;
        res->ngcap = ng;
        if (ng == 0) return SDCB200_OK;
        Scratch slots, sortkeys, order;
        SDC_TRY(slots.alloc((size_t)ng * 8, s));
        SDC_TRY(sortkeys.alloc((size_t)ng * 8, s));
        SDC_TRY(order.alloc((size_t)ng * 8, s));
        const int cgrid = (int)(ceil_div((int64_t)per, 256) < (int64_t)sms * 8 ? ceil_div((int64_t)per, 256) : (int64_t)sms * 8);
        gb_compact_kernel<<<cgrid, 256, 0, s>>>(t, slots.as<unsigned long long>(), sortkeys.as<unsigned long long>(),
                                                sort ? 0 : 1, dmeta);
        SDC_CHECK_LAUNCH();
        SDC_TRY(argsort_device(sort ? key_dtype : SDCB200_U64, sortkeys.p, ng, true, order.as<int64_t>(), s));
        SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->keys), (size_t)ng * 8, s));
        SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->vals), (size_t)ng * 8 * (nagg * cols.ncols + 1), s));
        const int fgrid = (int)(ceil_div(ng, 256) < (int64_t)sms * 4 ? ceil_div(ng, 256) : (int64_t)sms * 4);
        gb_finalize_kernel<<<fgrid, 256, 0, s>>>(t, cols.ncols, is_f64_mask, agg_mask, ddof, slots.as<unsigned long long>(),
                                                 order.as<int64_t>(), ng, res->keys, res->vals);
        SDC_CHECK_LAUNCH();
        SDC_CUDA_TRY(cudaStreamSynchronize(s));   // table scratch is released when `mem` goes out of scope
        return SDCB200_OK;
    }
    return SDCB200_ERR_CAPACITY;
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int32_t sdcb200_groupby(int32_t key_dtype, const void* keys, int64_t n, int32_t ncols, const void* const* cols,
                        const int32_t* col_dtypes, uint32_t agg_mask, int32_t sort, int64_t ddof,
                        int64_t expected_groups, void* stream, int64_t* handle_out, int64_t* ngroups_out) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(n >= 0 && ncols >= 0 && ncols <= kMaxCols, "n / ncols");
    SDC_ARG_CHECK(agg_mask != 0 && (agg_mask & ~0x7fu) == 0, "agg mask");
    SDC_ARG_CHECK(handle_out != nullptr && ngroups_out != nullptr, "null out pointer");
    SDC_ARG_CHECK(n == 0 || keys != nullptr, "null keys");
    Cols c;
    memset(&c, 0, sizeof(c));
    c.ncols = ncols;
    for (int i = 0; i < ncols; ++i) {
        SDC_ARG_CHECK(col_dtypes[i] == SDCB200_I64 || col_dtypes[i] == SDCB200_F64, "column dtype must be I64 or F64");
        SDC_ARG_CHECK(n == 0 || cols[i] != nullptr, "null column");
        c.data[i] = cols[i];
        c.is_f64[i] = col_dtypes[i] == SDCB200_F64;
    }
    GroupbyResult* res = new GroupbyResult();
    res->key_dtype = key_dtype;
    res->ncols = ncols;
    res->agg_mask = agg_mask;
    int32_t rc = SDCB200_OK;
    if (n > 0) {
        rc = key_dtype == SDCB200_F64
                 ? run_groupby<true>(keys, n, c, agg_mask, sort != 0, ddof, key_dtype, expected_groups, s, res)
                 : run_groupby<false>(keys, n, c, agg_mask, sort != 0, ddof, key_dtype, expected_groups, s, res);
    }
    if (rc != SDCB200_OK) {
        free_result(res);
        return rc;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    const int64_t h = g_next_handle++;
    g_results[h] = res;
    *handle_out = h;
    *ngroups_out = res->ngroups;
    return SDCB200_OK;
}

static GroupbyResult* find_result(int64_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_results.find(h);
    return it == g_results.end() ? nullptr : it->second;
}

int32_t sdcb200_groupby_keys(int64_t handle, void* dst, int32_t dst_is_host, void* stream) {
    GroupbyResult* r = find_result(handle);
    if (!r) { set_error("groupby: unknown handle %lld", (long long)handle); return SDCB200_ERR_HANDLE; }
    if (r->ngroups == 0) return SDCB200_OK;
    SDC_CUDA_TRY(cudaMemcpyAsync(dst, r->keys, (size_t)r->ngroups * 8,
                                 dst_is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (dst_is_host) SDC_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return SDCB200_OK;
}

int32_t sdcb200_groupby_values(int64_t handle, int32_t col, uint32_t agg, void* dst, int32_t dst_is_host, void* stream) {
    GroupbyResult* r = find_result(handle);
    if (!r) { set_error("groupby: unknown handle %lld", (long long)handle); return SDCB200_ERR_HANDLE; }
    SDC_ARG_CHECK(agg != 0 && (agg & (agg - 1)) == 0 && (r->agg_mask & agg), "aggregate was not requested");
    SDC_ARG_CHECK(col >= 0 && col < r->ncols, "column index");
    if (r->ngroups == 0) return SDCB200_OK;
    const int rank = __builtin_popcount(r->agg_mask & (agg - 1));
    SDC_CUDA_TRY(cudaMemcpyAsync(dst, r->vals + ((int64_t)rank * r->ncols + col) * r->ngcap, (size_t)r->ngroups * 8,
                                 dst_is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (dst_is_host) SDC_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return SDCB200_OK;
}

int32_t sdcb200_groupby_free(int64_t handle) {
    GroupbyResult* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_results.find(handle);
        if (it == g_results.end()) { set_error("groupby: unknown handle %lld", (long long)handle); return SDCB200_ERR_HANDLE; }
        r = it->second;
        g_results.erase(it);
    }
    free_result(r);
    return SDCB200_OK;
}

}  // extern "C"
