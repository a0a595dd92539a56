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
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>
#include <unordered_map>
#include <vector>

#include "radix_sort.cuh"
#include "scan.cuh"

namespace sdcb200 {
namespace {

// empty-slot marker: a NaN payload, never a canonical float key; an int64 key equal to it uses the spare slot
__host__ __device__ constexpr unsigned long long kEmptyKey() { return 0x7ff8dead0000beefull; }
constexpr int kMaxCols = 32;
constexpr int kGbThreads = 256;

enum : unsigned { AGG_SUM = 1, AGG_MEAN = 2, AGG_MIN = 4, AGG_MAX = 8, AGG_COUNT = 16, AGG_VAR = 32, AGG_STD = 64,
                  AGG_PROD = 128, AGG_MEDIAN = 256, AGG_LAST = AGG_MEDIAN };
enum { NEED_SUM = 1, NEED_CNT = 2, NEED_MIN = 4, NEED_MAX = 8, NEED_FIRST = 16, NEED_PROD = 32 };

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

// The global table.  ONE key array of cap + 1 slots (slot cap = the spare slot for key == kEmpty) and
// `nrep` replicas of every value array: the shared-memory path flushes CTA b into replica b % nrep so
// that the flush atomics of hundreds of CTAs do not serialise on the same few addresses (the key of a
// group is claimed once and only read afterwards, so the key array needs no replicas).  The last CTA
// to finish folds replicas 1.. into replica 0.  Value index = replica * (cap + 1) + slot.
struct Table {
    unsigned long long* keys;       // [cap + 1]
    unsigned long long* first;      // [nrep][cap + 1] first row of the group (NEED_FIRST)
    // per column c (nullptr when not needed), each [nrep][cap + 1]
    void* sum[kMaxCols];            // double or long long
    unsigned long long* cnt[kMaxCols];
    unsigned long long* mn[kMaxCols];   // ordered u64
    unsigned long long* mx[kMaxCols];
    double* ssd[kMaxCols];          // second pass: sum of squared deviations (replica 0 only)
    void* prod[kMaxCols];           // sort-and-segment path only: product of the non-NaN values (double or long long)
    double* med[kMaxCols];          // sort-and-segment path only: median of the non-NaN values
    unsigned long long* start;      // sort-and-segment path only: first sorted position of the group
    unsigned long long cap;         // power of two
    unsigned long long limit;       // max occupied slots before "overflow"
    unsigned long long* occupied;   // device counter
    int* overflow;                  // device flag: 1 = table too small, 2 = a key fell outside the direct window
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

// find-or-insert; returns the slot, or ~0 when the table overflowed
__device__ __forceinline__ unsigned long long global_upsert(const Table& t, unsigned long long key) {
    if (key == kEmptyKey()) {
        if (atomicCAS(&t.keys[t.cap], kEmptyKey(), key + 1) == kEmptyKey()) atomicAdd(t.occupied, 1ull);
        return t.cap;
    }
    unsigned long long slot = mix64(key) & (t.cap - 1);
    for (unsigned long long probes = 0; probes <= t.cap; ++probes) {
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&t.keys[slot]);
        if (cur == key) return slot;
        if (cur == kEmptyKey()) {
            const unsigned long long old = atomicCAS(&t.keys[slot], kEmptyKey(), key);
            if (old == kEmptyKey()) {
                if (atomicAdd(t.occupied, 1ull) + 1 > t.limit) { *(volatile int*)t.overflow = 1; return ~0ull; }
                return slot;
            }
            if (old == key) return slot;
        }
        slot = (slot + 1) & (t.cap - 1);
        // a table that filled up is abandoned by the host anyway: without this test every row still in flight walks
        // all `cap` slots of it (200 M rows over 1 M sparse keys and no hint: 94 ms in the 8 K-slot first attempt)
        if ((probes & 63ull) == 63ull && *(volatile int*)t.overflow) return ~0ull;
    }
    *(volatile int*)t.overflow = 1;
    return ~0ull;
}

// read-only lookup (second pass); the key is known to be present.  drange != 0: the table is
// direct-addressed (slot = key - dlo), as the shared-memory path left it for dense int64 keys.
__device__ __forceinline__ unsigned long long global_find(const Table& t, unsigned long long key, unsigned long long dlo,
                                                          unsigned long long drange) {
    if (drange) return key - dlo;
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

// ---- tile loads: a CTA of THREADS threads works on THREADS x 4 rows; thread t holds rows
// base + 2t, 2t+1, 2 THREADS + 2t, 2 THREADS + 2t + 1, so each column costs two coalesced 128-bit
// loads per thread and all loads of a tile are in flight together.
constexpr int kRows = 4;
constexpr int kGbTile = kGbThreads * kRows;

template <int THREADS = kGbThreads>
__device__ __forceinline__ int64_t tile_row(int64_t base, int tid, int e) {
    return base + (e >> 1) * (2 * THREADS) + 2 * tid + (e & 1);
}

template <int THREADS = kGbThreads>
__device__ __forceinline__ void load4(const void* col, int64_t base, int64_t n, int tid, bool vec,
                                      unsigned long long v[kRows]) {
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(col);
    if (vec && base + THREADS * kRows <= n) {
        // streamed once: first in line for L2 eviction, so the table lines the atomics work on stay resident
        const uint64_t pol = l2_policy_evict_first();
        const int4 a = ld_hint_v4(p + base + 2 * tid, pol);
        const int4 b = ld_hint_v4(p + base + 2 * THREADS + 2 * tid, pol);
        v[0] = (unsigned long long)(unsigned)a.x | ((unsigned long long)(unsigned)a.y << 32);
        v[1] = (unsigned long long)(unsigned)a.z | ((unsigned long long)(unsigned)a.w << 32);
        v[2] = (unsigned long long)(unsigned)b.x | ((unsigned long long)(unsigned)b.y << 32);
        v[3] = (unsigned long long)(unsigned)b.z | ((unsigned long long)(unsigned)b.w << 32);
    } else {
#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            const int64_t r = tile_row<THREADS>(base, tid, e);
            v[e] = r < n ? p[r] : 0ull;
        }
    }
}

// ------------------------------------------------------------------ high-cardinality path
// strided sample of an int64 key column by the whole CTA -> window [lo, lo + width) centred on the sampled
// keys when they span fewer than `width` values (range 0 otherwise).  Every CTA reads the same samples, so
// all CTAs agree.  s_red: 2 * THREADS / 32 words of shared scratch; the result is valid for thread 0 only.
template <int THREADS, int PER>
__device__ __forceinline__ void sample_window(const long long* __restrict__ kp, int64_t n, unsigned long long width,
                                              unsigned long long* s_red, unsigned long long& lo, unsigned long long& range,
                                              long long (&kv)[PER], bool tight = false) {
    constexpr int WARPS = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t total = (int64_t)THREADS * PER;
    const int64_t stride = n > total ? n / total : 1;
#pragma unroll
    for (int e = 0; e < PER; ++e) {        // independent loads, all in flight together
        const int64_t r = ((int64_t)e * THREADS + tid) * stride;
        kv[e] = r < n ? __ldg(kp + r) : __ldg(kp);
    }
    const unsigned long long top = 0x8000000000000000ull;   // order-preserving unsigned image: no overflow
    unsigned long long mn = ~0ull, mx = 0ull;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const unsigned long long u = (unsigned long long)kv[e] ^ top;
        mn = u < mn ? u : mn;
        mx = u > mx ? u : mx;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, d), b = __shfl_xor_sync(0xffffffffu, mx, d);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if (lane == 0) { s_red[warp] = mn; s_red[WARPS + warp] = mx; }
    __syncthreads();
    lo = 0;
    range = 0;
    if (tid == 0) {
        for (int w = 1; w < WARPS; ++w) {
            mn = s_red[w] < mn ? s_red[w] : mn;
            mx = s_red[WARPS + w] > mx ? s_red[WARPS + w] : mx;
        }
        // tight: the smallest power of two that leaves the sampled span as much room again (the caller replicates the
        // window across the rest of the table)
        if (tight && mx - mn < width) {
            unsigned long long w = 64;
            while (w < width && w < 2 * (mx - mn + 1)) w <<= 1;
            width = w < width ? w : width;
        }
        if (mx - mn < width) {
            unsigned long long shift = (width - (mx - mn) - 1) / 2;
            if (shift > mn) shift = mn;
            unsigned long long ulo = mn - shift;
            if (ulo > ~0ull - width) ulo = ~0ull - width;
            lo = ulo ^ top;
            range = width;
        }
    }
}

// key modes of the int64 path
enum { KEYS_PLAIN = 0, KEYS_GID = 1 };      // GID: dense group ids in [0, range), negative = skip the row

struct DirectHint {
    int mode;                       // KEYS_PLAIN: sample the keys (allow != 0); KEYS_GID: window [0, range)
    int allow;                      // 0: hashing only (a previous attempt saw a key outside its window)
    unsigned long long range;       // KEYS_GID: the id range
    unsigned long long* out;        // [2] device: the window the kernel used (lo, range), read back by the host
};

template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_global_kernel(const void* __restrict__ keys, int64_t n, Cols cols,
                                                               Table t, int need, DirectHint dh) {
    __shared__ unsigned long long s_red[2 * kGbThreads / 32];
    __shared__ unsigned long long s_dlo, s_drange;
    const int tid = threadIdx.x;
    // dense int64 keys: the table is direct-addressed (slot = key - lo): no hashing, no CAS, presence = the key word
    if (!KEY_F64 && dh.mode == KEYS_GID) {
        if (tid == 0) { s_dlo = 0ull; s_drange = dh.range; }
    } else if (!KEY_F64 && dh.allow) {
        unsigned long long lo, range;
        long long kv_unused[4];
        sample_window<kGbThreads, 4>(reinterpret_cast<const long long*>(keys), n, t.cap, s_red, lo, range, kv_unused);
        if (tid == 0) { s_dlo = lo; s_drange = range; }
    } else if (tid == 0) {
        s_dlo = 0ull;
        s_drange = 0ull;
    }
    __syncthreads();
    const unsigned long long dlo = s_dlo, drange = s_drange;
    if (blockIdx.x == 0 && tid == 0) { dh.out[0] = dlo; dh.out[1] = drange; }
    const bool skip_neg = !KEY_F64 && dh.mode == KEYS_GID;
    const int64_t ntiles = (n + kGbTile - 1) / kGbTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // probe loops end at different times per lane; a warp that stays split would issue the next tile's
        // 128-bit loads as several partial requests
        __syncwarp();
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
            __syncwarp();
            if (!(tile_row(base, tid, e) < n && canon_key<KEY_F64>(kb[e], k))) continue;
            if (skip_neg && (long long)k < 0) continue;
            if (drange) {
                const unsigned long long d = k - dlo;
                if (d < drange && k != kEmptyKey()) {
                    g[e] = d;
                    if (t.keys[d] != k) t.keys[d] = k;              // every writer stores the same value
                } else {
                    *(volatile int*)t.overflow = 2;                 // outside the window: rerun with hashing
                }
            } else {
                g[e] = global_upsert(t, k);
            }
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
        case AGG_PROD:
            return reinterpret_cast<const unsigned long long*>(t.prod[c])[slot];
        case AGG_MEDIAN:
            return (unsigned long long)__double_as_longlong(t.med[c][slot]);
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
    for (unsigned bit = 1; bit <= AGG_LAST; bit <<= 1) {
        if (!(agg_mask & bit)) continue;
        for (int c = 0; c < ncols; ++c)
            out_vals[((int64_t)rank * ncols + c) * ngcap + j] = agg_value(t, (is_f64_mask >> c) & 1u, c, bit, ddof, slot);
        ++rank;
    }
}

// ------------------------------------------------------------------ low-cardinality path
// ONE kernel.  Every CTA (512 threads, two per SM, persistent over tiles of 2048 rows):
//   1. samples the key column (int64 keys): if the sampled keys span fewer values than the shared
//      table has slots, the table is DIRECT-ADDRESSED (slot = key - lo: no hashing, no probing, and
//      slot order = key order); every CTA reads the same samples, so all agree on the window.
//      A later key outside the window raises overflow = 2 and the host reruns with hashing.
//   2. aggregates its rows into a private shared-memory table (SoA):
//        keys[SCAP] u64 | first[SCAP] u64 | per column: sum[SCAP] 8B, mn[SCAP] u64, mx[SCAP] u64 | cnt[SCAP] u32 per column
//      (only the arrays `need` asks for are carved out).  Hashing mode: rows that cannot get a slot
//      within kSmemProbes probes go to the global table directly.
//   3. flushes its occupied slots into replica (blockIdx % nrep) of the small global table.
//   4. takes a ticket; the LAST CTA folds the replicas into replica 0, compacts the groups, orders
//      them (direct + sort=True: already in key order, a block scan; otherwise a bitonic sort of
//      (key | first row, slot) in shared memory) and writes every aggregate -- no second launch.
// HBM traffic = keys + values read once.
constexpr int kSmThreads = 512;                 // two CTAs per SM: two private tables halve the CAS contention
constexpr int kSmemProbes = 16;
constexpr int kLeanProbes = 64;                  // hashed lean loop: a longer walk is cheaper than the rerun it would cause
constexpr int kSamplePerThread = 4;              // 2048 strided samples per CTA
constexpr int kSmallReplicas = 8;                // replicas of the small global table's value arrays

struct SmemLayout {
    int scap;
    int off_first;          // byte offsets from the table base; -1 = absent
    int off_sum[kMaxCols], off_cnt[kMaxCols], off_mn[kMaxCols], off_mx[kMaxCols];
    int bytes;              // table bytes
    int alloc_bytes;        // dynamic shared memory of the launch (table or the finalize's sort arrays)
};

struct SmallFin {
    unsigned long long* out_keys;
    unsigned long long* out_vals;
    int64_t ngcap;
    unsigned agg_mask, is_f64_mask;
    long long ddof;
    int by_first, key_f64;
    int write_out;                  // 0: only fold the replicas (var / std continue on the host path)
    int rf_max;                     // lean loop: most lane-indexed copies of the direct window inside the shared table (1 = off)
    int hot;                        // lean loop: keep the most frequent sampled keys' aggregates in registers
    int lean_hash;                  // lean instantiation: keys without a direct window take the hashed lean loop (2b)
    DirectHint dh;                  // dh.out = meta + 2
    unsigned long long* meta;       // [0] groups written, [1] CTAs done, [2] direct lo, [3] direct range
    unsigned long long* host_meta;  // mapped pinned host copy of the 256-byte control block, written by the last CTA (or null)
};

// SPEC >= 0: compile-time `need` mask for ONE float64 value column (the common df.groupby(k).agg() shapes get a
// tight loop); SPEC < 0: any columns / aggregates, decided at run time.
// ---- LEAN instantiations of gb_smem_kernel (one float64 column, direct-addressed slots): the per-row work of the
// generic loop is 92 instructions of which a quarter is the float64 add itself (ncu, profiles/r1_groupby_100m_*) and
// the kernel is issue-bound (62 % issue slots busy).  The lean loop drops everything a direct window does not need
// (probe loop, global-table fallback, slot arrays, overflow polling) and addresses shared memory through explicit
// 32-bit shared addresses: with generic pointers the compiler rebuilt the shared window base (five uniform
// instructions) in front of every access.  `red.shared.add.f64` expands to the same load / add / CAS-spin loop as
// atomicAdd(double*).
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x) {      // keeps an address in a register instead of
    asm volatile("" : "+r"(x));                                   // letting the compiler rematerialise it
    return x;
}
__device__ __forceinline__ void reds_add_f64(uint32_t a, double v) { asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void reds_add_u32(uint32_t a, unsigned v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_min_u64(uint32_t a, unsigned long long v) { asm volatile("red.shared.min.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void reds_max_u64(uint32_t a, unsigned long long v) { asm volatile("red.shared.max.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_zero_u32(uint32_t a) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(0u) : "memory"); }

template <bool KEY_F64, int SPEC, bool LEAN = false>
__global__ void __launch_bounds__(kSmThreads, 2) gb_smem_kernel(const void* __restrict__ keys, int64_t n, Cols cols_rt, Table t,
                                                                int need_rt, SmemLayout lay, SmallFin fin) {
    static_assert(!LEAN || SPEC >= 0, "lean loops: one float64 value column");
    const int need = SPEC >= 0 ? SPEC : need_rt;
    struct ColsView {
        const Cols& c;
        __device__ __forceinline__ int ncols() const { return SPEC >= 0 ? 1 : c.ncols; }
        __device__ __forceinline__ bool is_f64(int i) const { return SPEC >= 0 ? true : c.is_f64[i] != 0; }
    };
    const ColsView cv{cols_rt};
    const Cols& cols = cols_rt;
    constexpr int kSmTile = kSmThreads * kRows;
    constexpr int kSmWarps = kSmThreads / 32;
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ unsigned long long s_red[2 * kSmWarps];
    __shared__ unsigned long long s_dlo, s_drange;
    __shared__ unsigned s_counter;
    __shared__ int s_last;
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sm);
    unsigned long long* sfirst = reinterpret_cast<unsigned long long*>(sm + (lay.off_first >= 0 ? lay.off_first : 0));
    const int scap = lay.scap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rep = blockIdx.x % t.nrep;
    const unsigned long long per = t.cap + 1;

    // the first tile's loads go out before anything else; inside the loop the next tile is always in flight
    const int64_t ntiles = (n + kSmTile - 1) / kSmTile;
    unsigned long long kbn[kRows], v0n[kRows];
    if ((int64_t)blockIdx.x < ntiles) {
        load4<kSmThreads>(keys, (int64_t)blockIdx.x * kSmTile, n, tid, cols.vec_ok, kbn);
        if (cv.ncols() > 0) load4<kSmThreads>(cols.data[0], (int64_t)blockIdx.x * kSmTile, n, tid, cols.vec_ok, v0n);
    }

    // ---- 1. key sample -> direct window: a TIGHT one (a power of two that leaves the sampled span as much room
    //      again) so that the rest of the table can hold further copies of it -- see 2 (LEAN)
    const unsigned long long wmax = (unsigned long long)scap < t.cap ? (unsigned long long)scap : t.cap;
    const bool tight = fin.rf_max > 1;
    long long kvs[kSamplePerThread];
    bool sampled = false;
    if (!KEY_F64 && fin.dh.mode == KEYS_GID) {
        if (tid == 0) {
            unsigned long long w = wmax;
            if (tight) { w = 64; while (w < wmax && w < fin.dh.range) w <<= 1; }
            s_dlo = 0ull;
            s_drange = fin.dh.range <= w ? w : 0ull;
        }
        if (LEAN && fin.hot) {             // group ids need no window, but the hot-key search wants the same key samples
            const long long* kp = reinterpret_cast<const long long*>(keys);
            const int64_t total = (int64_t)kSmThreads * kSamplePerThread;
            const int64_t stride = n > total ? n / total : 1;
#pragma unroll
            for (int e = 0; e < kSamplePerThread; ++e) {
                const int64_t r = ((int64_t)e * kSmThreads + tid) * stride;
                kvs[e] = r < n ? __ldg(kp + r) : -1;        // a negative id is a skipped row
            }
            sampled = true;
        }
    } else if (!KEY_F64 && fin.dh.allow) {
        unsigned long long lo, range;
        sample_window<kSmThreads, kSamplePerThread>(reinterpret_cast<const long long*>(keys), n, wmax, s_red, lo, range, kvs, tight);
        if (tid == 0) { s_dlo = lo; s_drange = range; }
        sampled = true;
    } else if (tid == 0) {
        s_dlo = 0ull;
        s_drange = 0ull;
    }
    // ---- 1b (LEAN). hot keys: a histogram of the samples in the (not yet initialised) table memory, then the kHot
    //      most frequent slots that hold at least 2 % of the samples each.  Their rows skip the float64 CAS loop: a
    //      thread keeps their sums and counts in registers (see row_hot below), so a skewed key distribution does not
    //      become a CAS retry storm on a few slots.
    constexpr int kHot = 4;
    uint32_t hot[kHot];
#pragma unroll
    for (int h = 0; h < kHot; ++h) hot[h] = 0xffffffffu;
    int nhot = 0;
    if (LEAN && fin.hot && sampled) {
        __syncthreads();                                   // s_dlo / s_drange visible; s_red free again
        const unsigned long long hlo = s_dlo, hrange = s_drange;
        if (hrange) {
            unsigned* hist = reinterpret_cast<unsigned*>(sm);
            unsigned* s_red32 = reinterpret_cast<unsigned*>(s_red);
            for (unsigned i = tid; i < (unsigned)hrange; i += kSmThreads) hist[i] = 0u;
            __syncthreads();
#pragma unroll
            for (int e = 0; e < kSamplePerThread; ++e) {
                const unsigned long long d = (unsigned long long)kvs[e] - hlo;
                if (d < hrange) atomicAdd(hist + (unsigned)d, 1u);
            }
            __syncthreads();
            constexpr unsigned kHotMin = (unsigned)(kSmThreads * kSamplePerThread) / 50u;
#pragma unroll
            for (int h = 0; h < kHot; ++h) {
                unsigned best = 0u;
                for (unsigned i = tid; i < (unsigned)hrange; i += kSmThreads) {
                    const unsigned pk = (hist[i] << 16) | i;
                    best = pk > best ? pk : best;
                }
                best = __reduce_max_sync(0xffffffffu, best);
                if (lane == 0) s_red32[warp] = best;
                __syncthreads();
                best = 0u;
#pragma unroll
                for (int w = 0; w < kSmWarps; ++w) best = s_red32[w] > best ? s_red32[w] : best;
                if ((best >> 16) < kHotMin) break;         // block-uniform: no further slot reaches 2 % either (uniform keys
                                                           // leave after the first round instead of searching four times)
                hot[h] = best & 0xffffu;
                nhot = h + 1;
                if ((int)((best & 0xffffu) % kSmThreads) == tid) hist[best & 0xffffu] = 0u;
                __syncthreads();
            }
            __syncthreads();                               // every thread has read s_red32 / hist before they are reused
        }
    }
    if (blockIdx.x == 0 && tid == 0) { fin.dh.out[0] = s_dlo; fin.dh.out[1] = s_drange; }
    const bool skip_neg = !KEY_F64 && fin.dh.mode == KEYS_GID;
    // The lean loop does not store a presence mark per row: a slot is occupied when its count is non-zero or, with no
    // count array (sum only), when its sum no longer holds the initial -0.0 (x + -0.0 == x, and a sum can only come back
    // to -0.0 from terms that are all -0.0); rows that leave both untouched (NaN values, zeros) store the mark.
    constexpr bool kCntPresence = LEAN && (SPEC & NEED_CNT) != 0;
    constexpr bool kSumOnlyLean = LEAN && !kCntPresence;
    constexpr unsigned long long kNegZero = 0x8000000000000000ull;
    for (int i = tid; i < scap; i += kSmThreads) {
        skeys[i] = kEmptyKey();
        if (need & NEED_FIRST) sfirst[i] = ~0ull;
        for (int c = 0; c < cv.ncols(); ++c) {
            if (need & NEED_SUM) reinterpret_cast<unsigned long long*>(sm + lay.off_sum[c])[i] = kSumOnlyLean ? kNegZero : 0ull;
            if (need & NEED_CNT) reinterpret_cast<unsigned*>(sm + lay.off_cnt[c])[i] = 0u;
            if (need & NEED_MIN) reinterpret_cast<unsigned long long*>(sm + lay.off_mn[c])[i] = ~0ull;
            if (need & NEED_MAX) reinterpret_cast<unsigned long long*>(sm + lay.off_mx[c])[i] = 0ull;
        }
    }
    __syncthreads();
    const unsigned long long dlo = s_dlo, drange = s_drange;

    // ---- 2 (LEAN). aggregate, direct-addressed slots only.  The window of `drange` slots is copied rf times across the
    //      table and a row uses copy (lane % rf): rows of one warp instruction that share a key no longer share an
    //      address (with few groups rf reaches 32 and a warp never collides with itself); the flush adds the copies up.
    const unsigned rf = !drange ? 1u : [&] {
        unsigned r = (unsigned)((unsigned long long)scap / drange);
        r = r > 32u ? 32u : r;
        return r > (unsigned)fin.rf_max ? (unsigned)(fin.rf_max < 1 ? 1 : fin.rf_max) : r;
    }();
    if (LEAN && drange) {
        const uint32_t copy = ((uint32_t)lane & (rf - 1u)) * (uint32_t)drange;
        const uint32_t keys_a = opaque_u32(smem_u32(skeys));
        const uint32_t sum_a = opaque_u32(smem_u32(sm + lay.off_sum[0]) + copy * 8u);
        const uint32_t cnt_a = opaque_u32(smem_u32(sm + lay.off_cnt[0]) + copy * 4u);
        const uint32_t mn_a = opaque_u32(smem_u32(sm + lay.off_mn[0]) + copy * 8u);
        const uint32_t mx_a = opaque_u32(smem_u32(sm + lay.off_mx[0]) + copy * 8u);
        auto accumulate = [&](uint32_t d, double s, unsigned c, unsigned long long mn, unsigned long long mx, bool mark = false) {
            const uint32_t o8 = d * 8u;
            // explicit presence (copy 0 only; the high word of the slot's key leaves the empty pattern) for the rows the
            // count / sum cannot vouch for
            if (mark || c == 0u || (kSumOnlyLean && s == 0.0)) sts_zero_u32(keys_a + o8 + 4u);
            if (c) {
                if (need & NEED_SUM) reds_add_f64(sum_a + o8, s);
                if (need & NEED_CNT) reds_add_u32(cnt_a + d * 4u, c);
                // 64-bit shared min / max are CAS loops too, and a group's extremes settle after a few rows: a plain
                // load first (a stale value only costs an atomic that changes nothing)
                if (need & NEED_MIN) { if (mn < lds_u64(mn_a + o8)) reds_min_u64(mn_a + o8, mn); }
                if (need & NEED_MAX) { if (mx > lds_u64(mx_a + o8)) reds_max_u64(mx_a + o8, mx); }
            }
        };
        auto row = [&](unsigned long long k, unsigned long long vbits) {
            const unsigned long long d = k - dlo;
            if (d < drange && k != kEmptyKey()) {
                const double v = __longlong_as_double((long long)vbits);
                const unsigned long long o = (need & (NEED_MIN | NEED_MAX)) ? f64_to_ordered(v) : 0ull;
                accumulate((uint32_t)d, v, v == v ? 1u : 0u, o, o);
            } else if (!(skip_neg && (long long)k < 0)) {
                *(volatile int*)t.overflow = 2;             // outside the sampled window: rerun with hashing
            }
        };
        auto tiles = [&](auto&& full_row) {
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int64_t base = tile * kSmTile;
                unsigned long long kb[kRows], v0[kRows];
#pragma unroll
                for (int e = 0; e < kRows; ++e) { kb[e] = kbn[e]; v0[e] = v0n[e]; }
                if (tile + gridDim.x < ntiles) {
                    const int64_t nbase = (tile + gridDim.x) * kSmTile;
                    load4<kSmThreads>(keys, nbase, n, tid, cols.vec_ok, kbn);
                    load4<kSmThreads>(cols.data[0], nbase, n, tid, cols.vec_ok, v0n);
                }
                if (base + kSmTile <= n) {
#pragma unroll
                    for (int e = 0; e < kRows; ++e) full_row(kb[e], v0[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < kRows; ++e)
                        if (tile_row<kSmThreads>(base, tid, e) < n) row(kb[e], v0[e]);
                }
            }
        };
        if (!nhot) {
            tiles(row);
        } else {
            // hot keys: sum and count in registers, folded across the warp after the last tile; min / max go through
            // the shared slot, whose extremes settle after a few rows (a broadcast load, rarely an atomic)
            double hs[kHot];
            unsigned hc[kHot], hseen = 0u;
#pragma unroll
            for (int h = 0; h < kHot; ++h) { hs[h] = 0.0; hc[h] = 0u; }
            auto row_hot = [&](unsigned long long k, unsigned long long vbits) {
                const unsigned long long d = k - dlo;
                if (d < drange && k != kEmptyKey()) {
                    const double v = __longlong_as_double((long long)vbits);
                    const bool vv = v == v;
                    const unsigned long long o = (need & (NEED_MIN | NEED_MAX)) ? f64_to_ordered(v) : 0ull;
                    bool any = false;
#pragma unroll
                    for (int h = 0; h < kHot; ++h) {
                        const bool m = (uint32_t)d == hot[h];
                        const bool mv = m && vv;
                        any |= m;
                        hseen |= m ? (1u << h) : 0u;
                        if (need & NEED_SUM) hs[h] += mv ? v : 0.0;
                        hc[h] += mv ? 1u : 0u;
                    }
                    if (!any) accumulate((uint32_t)d, v, vv ? 1u : 0u, o, o);
                    else if ((need & (NEED_MIN | NEED_MAX)) && vv) {
                        const uint32_t o8 = (uint32_t)d * 8u;
                        if (need & NEED_MIN) { if (o < lds_u64(mn_a + o8)) reds_min_u64(mn_a + o8, o); }
                        if (need & NEED_MAX) { if (o > lds_u64(mx_a + o8)) reds_max_u64(mx_a + o8, o); }
                    }
                } else if (!(skip_neg && (long long)k < 0)) {
                    *(volatile int*)t.overflow = 2;
                }
            };
            tiles(row_hot);
#pragma unroll
            for (int h = 0; h < kHot; ++h) {
                const unsigned seen = __any_sync(0xffffffffu, (hseen >> h) & 1u);
                double sv = hs[h];
#pragma unroll
                for (int x = 16; x > 0; x >>= 1)
                    if (need & NEED_SUM) sv += __shfl_xor_sync(0xffffffffu, sv, x);
                const unsigned cn = __reduce_add_sync(0xffffffffu, hc[h]);
                // min / max are already in the slot: the identities leave them alone
                if (lane == 0 && seen) accumulate(hot[h], sv, cn, ~0ull, 0ull, true);
            }
        }
    } else if (LEAN && fin.lean_hash) {
        // ---- 2b (LEAN, hashed). Keys with no dense window -- sparse int64 ids, float64 keys, the 8-byte words of a
        //      fixed-width string column: the slot comes from a probe of the CTA's shared key table (each lane walks its
        //      own chain; at the load factors the layout allows one or two steps), the aggregates then take the lean
        //      loop's explicit shared addresses.  There is no global-table fallback inside the loop: a key that finds
        //      no slot within kLeanProbes steps (more groups than the table holds) raises overflow = 3 and the host
        //      reruns with the generic loop below.  The generic loop costs 2.2 ms per 100 M rows and five aggregates
        //      over 1 K sparse keys, the direct lean loop 0.61.
        const uint32_t keys_a = opaque_u32(smem_u32(skeys));
        const uint32_t sum_a = opaque_u32(smem_u32(sm + lay.off_sum[0]));
        const uint32_t cnt_a = opaque_u32(smem_u32(sm + lay.off_cnt[0]));
        const uint32_t mn_a = opaque_u32(smem_u32(sm + lay.off_mn[0]));
        const uint32_t mx_a = opaque_u32(smem_u32(sm + lay.off_mx[0]));
        const uint32_t smask = (uint32_t)scap - 1u;
        // the rest of a probe whose first slot held another key (or nothing): walk on, claim an empty slot
        auto probe_on = [&](unsigned long long k, uint32_t s, unsigned long long cur) -> uint32_t {
            if (k == kEmptyKey()) return 0xffffffffu;                        // the marker itself cannot live in this table
#pragma unroll 1
            for (int p = 0; p < kLeanProbes; ++p) {
                if (cur == k) return s;
                if (cur == kEmptyKey()) {
                    const unsigned long long old = atomicCAS(&skeys[s], kEmptyKey(), k);
                    if (old == kEmptyKey() || old == k) return s;
                }
                s = (s + 1u) & smask;
                cur = lds_u64(keys_a + s * 8u);
            }
            return 0xffffffffu;
        };
        auto add_row = [&](uint32_t s, unsigned long long vbits) {
            const double v = __longlong_as_double((long long)vbits);
            if (v != v) return;                                               // skipna: the group exists, nothing to add
            const uint32_t o8 = s * 8u;
            if (need & NEED_SUM) reds_add_f64(sum_a + o8, v);
            if (need & NEED_CNT) reds_add_u32(cnt_a + s * 4u, 1u);
            if (need & (NEED_MIN | NEED_MAX)) {
                const unsigned long long o = f64_to_ordered(v);
                if (need & NEED_MIN) { if (o < lds_u64(mn_a + o8)) reds_min_u64(mn_a + o8, o); }
                if (need & NEED_MAX) { if (o > lds_u64(mx_a + o8)) reds_max_u64(mx_a + o8, o); }
            }
        };
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t base = tile * kSmTile;
            unsigned long long kb[kRows], v0[kRows];
#pragma unroll
            for (int e = 0; e < kRows; ++e) { kb[e] = kbn[e]; v0[e] = v0n[e]; }
            if (tile + gridDim.x < ntiles) {
                const int64_t nbase = (tile + gridDim.x) * kSmTile;
                load4<kSmThreads>(keys, nbase, n, tid, cols.vec_ok, kbn);
                load4<kSmThreads>(cols.data[0], nbase, n, tid, cols.vec_ok, v0n);
            }
            const int ovf = *(volatile int*)t.overflow;      // consumed at the end of the iteration
            // the four rows' first probes are independent loads, all in flight together; only a row whose first slot
            // is not its key walks on, and the warp is whole again before the (expensive) shared atomics
            const bool full = base + kSmTile <= n;
            bool valid[kRows];
            uint32_t sl[kRows];
            unsigned long long cur[kRows];
#pragma unroll
            for (int e = 0; e < kRows; ++e) {
                unsigned long long k = 0;
                valid[e] = (full || tile_row<kSmThreads>(base, tid, e) < n) && canon_key<KEY_F64>(kb[e], k);   // NaN key: row skipped
                kb[e] = k;
                sl[e] = (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> 40) & smask;
            }
#pragma unroll
            for (int e = 0; e < kRows; ++e) cur[e] = lds_u64(keys_a + sl[e] * 8u);
#pragma unroll
            for (int e = 0; e < kRows; ++e)
                if (valid[e] && (cur[e] != kb[e] || kb[e] == kEmptyKey())) sl[e] = probe_on(kb[e], sl[e], cur[e]);
            __syncwarp();
#pragma unroll
            for (int e = 0; e < kRows; ++e) {
                if (!valid[e]) continue;
                if (sl[e] == 0xffffffffu) *(volatile int*)t.overflow = 3;
                else add_row(sl[e], v0[e]);
            }
            if (ovf) break;
        }
    } else
    // ---- 2. aggregate (generic: hashed or direct slots, any columns; a LEAN instantiation lands here when the
    //      sampled keys do not fit a direct window and the hashed lean loop is off or gave up).  Direct slots use the
    //      lane-indexed copies as well.
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int gcopy = (int)(((unsigned)lane & (rf - 1u)) * (unsigned)drange);
        __syncwarp();
        const int64_t base = tile * kSmTile;
        unsigned long long kb[kRows], v0[kRows];
#pragma unroll
        for (int e = 0; e < kRows; ++e) { kb[e] = kbn[e]; v0[e] = v0n[e]; }
        if (tile + gridDim.x < ntiles) {
            const int64_t nbase = (tile + gridDim.x) * kSmTile;
            load4<kSmThreads>(keys, nbase, n, tid, cols.vec_ok, kbn);
            if (cv.ncols() > 0) load4<kSmThreads>(cols.data[0], nbase, n, tid, cols.vec_ok, v0n);
        }
        const int ovf = *(volatile int*)t.overflow;      // consumed at the end of the iteration
        int slot[kRows];                      // >= 0 shared slot, -1 row skipped, -2 row lives in the global table
        unsigned long long g[kRows];
#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            unsigned long long k = 0;
            bool valid = tile_row<kSmThreads>(base, tid, e) < n && canon_key<KEY_F64>(kb[e], k);
            if (skip_neg && (long long)k < 0) valid = false;
            int sl = -1;
            g[e] = ~0ull;
            if (drange) {
                const unsigned long long d = k - dlo;
                if (valid) {
                    if (d < drange && k != kEmptyKey()) {
                        sl = (int)d + gcopy;                // this lane's copy of the window
                        // presence marker: one native 32-bit shared atomic on the slot's key word (the key itself is
                        // dlo + slot); it cannot wrap back to the empty pattern within a CTA's rows
                        atomicAdd(reinterpret_cast<unsigned*>(skeys + sl), 1u);
                    } else {
                        *(volatile int*)t.overflow = 2;             // outside the sampled window: rerun with hashing
                    }
                }
            } else {
                bool pending = valid && k != kEmptyKey();
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
                if (valid && sl < 0) {      // shared table full around this key (or key == kEmpty): global table
                    const unsigned long long gs = global_upsert(t, k);
                    if (gs != ~0ull) { g[e] = (unsigned long long)rep * per + gs; sl = -2; }
                }
            }
            slot[e] = sl;
            __syncwarp();
        }
        if (need & NEED_FIRST) {
#pragma unroll
            for (int e = 0; e < kRows; ++e) {
                const unsigned long long row = (unsigned long long)tile_row<kSmThreads>(base, tid, e);
                if (slot[e] >= 0) { if (row < sfirst[slot[e]]) atomicMin(&sfirst[slot[e]], row); }
                else if (slot[e] == -2) atomicMin(&t.first[g[e]], row);
            }
            __syncwarp();
        }
        for (int c = 0; c < cv.ncols(); ++c) {
            if (c > 0) load4<kSmThreads>(cols.data[c], base, n, tid, cols.vec_ok, v0);
            const bool f = cv.is_f64(c);
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
                            const unsigned long long o = f64_to_ordered(v);      // extremes settle early: load, then atomic
                            if ((need & NEED_MIN) && o < *(volatile unsigned long long*)(smn + sl)) atomicMin(smn + sl, o);
                            if ((need & NEED_MAX) && o > *(volatile unsigned long long*)(smx + sl)) atomicMax(smx + sl, o);
                        }
                    } else {
                        if (need & NEED_SUM) atomicAdd(reinterpret_cast<unsigned long long*>(ssum) + sl, v0[e]);
                        if (need & NEED_CNT) atomicAdd(scnt + sl, 1u);
                        const unsigned long long o = i64_to_ordered((long long)v0[e]);
                        if ((need & NEED_MIN) && o < *(volatile unsigned long long*)(smn + sl)) atomicMin(smn + sl, o);
                        if ((need & NEED_MAX) && o > *(volatile unsigned long long*)(smx + sl)) atomicMax(smx + sl, o);
                    }
                }
                __syncwarp();
            }
        }
        if (ovf) break;
    }
    __syncthreads();

    // ---- 3. flush the occupied shared slots into this CTA's replica of the global value arrays
    if (LEAN && drange) {
        // presence lives in copy 0 of the key words; the value arrays hold rf copies of the window, added up here
        const unsigned S = (unsigned)drange;
        for (unsigned d = tid; d < S; d += kSmThreads) {
            bool present = skeys[d] != kEmptyKey();
            for (unsigned r = 0; r < rf && !present; ++r)
                present = kCntPresence ? reinterpret_cast<unsigned*>(sm + lay.off_cnt[0])[r * S + d] != 0u
                                       : reinterpret_cast<unsigned long long*>(sm + lay.off_sum[0])[r * S + d] != kNegZero;
            if (!present) continue;
            const unsigned long long k = dlo + (unsigned long long)d;
            if (*reinterpret_cast<volatile unsigned long long*>(&t.keys[d]) != k) t.keys[d] = k;   // same value from every CTA
            const unsigned long long g = (unsigned long long)rep * per + d;
            if (need & NEED_CNT) {
                unsigned cn = 0u;
                for (unsigned r = 0; r < rf; ++r) cn += reinterpret_cast<unsigned*>(sm + lay.off_cnt[0])[r * S + d];
                if (cn) atomicAdd(&t.cnt[0][g], (unsigned long long)cn);
            }
            if (need & NEED_SUM) {
                double v = 0.0;
                for (unsigned r = 0; r < rf; ++r) v += reinterpret_cast<double*>(sm + lay.off_sum[0])[r * S + d];
                if (v != 0.0) atomicAdd(reinterpret_cast<double*>(t.sum[0]) + g, v);
            }
            if (need & NEED_MIN) {
                unsigned long long m = ~0ull;
                for (unsigned r = 0; r < rf; ++r) { const unsigned long long o = reinterpret_cast<unsigned long long*>(sm + lay.off_mn[0])[r * S + d]; m = o < m ? o : m; }
                atomicMin(&t.mn[0][g], m);
            }
            if (need & NEED_MAX) {
                unsigned long long m = 0ull;
                for (unsigned r = 0; r < rf; ++r) { const unsigned long long o = reinterpret_cast<unsigned long long*>(sm + lay.off_mx[0])[r * S + d]; m = o > m ? o : m; }
                atomicMax(&t.mx[0][g], m);
            }
        }
    } else
    for (int i = tid; i < scap; i += kSmThreads) {
        unsigned long long k = skeys[i];
        if (k == kEmptyKey()) continue;
        unsigned long long gs;
        if (drange) {
            gs = (unsigned long long)i & (drange - 1ull);   // any of the window's copies (drange is a power of two)
            k = dlo + gs;
            if (*reinterpret_cast<volatile unsigned long long*>(&t.keys[gs]) != k) t.keys[gs] = k;   // same value from every CTA
        } else {
            gs = global_upsert(t, k);
            if (gs == ~0ull) continue;
        }
        const unsigned long long g = (unsigned long long)rep * per + gs;
        if (need & NEED_FIRST) atomicMin(&t.first[g], sfirst[i]);
        for (int c = 0; c < cv.ncols(); ++c) {
            if (need & NEED_CNT) {
                const unsigned cn = reinterpret_cast<unsigned*>(sm + lay.off_cnt[c])[i];
                if (cn) atomicAdd(&t.cnt[c][g], (unsigned long long)cn);
            }
            if (need & NEED_SUM) {
                if (cv.is_f64(c)) {
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

    // ---- 4. the last CTA to get here finishes the job
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned long long ticket = atomicAdd(&fin.meta[1], 1ull);
        s_last = ticket == (unsigned long long)gridDim.x - 1;
        s_counter = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the host reads the control block from mapped pinned memory once the stream is idle: no copy-engine round trip
    auto publish = [&]() {
        if (!fin.host_meta) return;
        __syncthreads();
        if (tid < 32) fin.host_meta[tid] = reinterpret_cast<volatile unsigned long long*>(t.occupied)[tid];
        __threadfence_system();
    };
    if (*(volatile int*)t.overflow) { publish(); return; }
    // a direct window only ever touches slots [0, drange): fold / compact those (without a hint the table has 8193 slots
    // per replica, the window of 1 K groups 2048)
    const unsigned nslots = drange ? (unsigned)(drange < per ? drange : per) : (unsigned)per;
    // fold replicas 1.. into replica 0: plain L2 loads (every other CTA is done), unconditional (an empty slot folds
    // its initial values), two slots per step so that 2 x nrep loads are in flight per thread
    if (t.nrep > 1) {
        auto comb = [](unsigned long long a, unsigned long long b, int mode) -> unsigned long long {
            if (mode == 0) return a + b;
            if (mode == 1) return (unsigned long long)__double_as_longlong(__longlong_as_double((long long)a) + __longlong_as_double((long long)b));
            if (mode == 2) return b < a ? b : a;
            return b > a ? b : a;
        };
        auto fold = [&](unsigned long long* arr, int mode /* 0 add u64, 1 add f64, 2 min, 3 max */) {
            for (unsigned i = tid; i < nslots; i += 2 * kSmThreads) {
                const unsigned i2 = i + kSmThreads;
                const bool two = i2 < nslots;
                unsigned long long v[kSmallReplicas], w[kSmallReplicas];
#pragma unroll
                for (int r = 0; r < kSmallReplicas; ++r) {
                    v[r] = __ldcg(arr + (r < t.nrep ? r : 0) * per + i);
                    w[r] = __ldcg(arr + (r < t.nrep ? r : 0) * per + (two ? i2 : i));
                }
                unsigned long long a = v[0], c = w[0];
#pragma unroll
                for (int r = 1; r < kSmallReplicas; ++r) {
                    if (r < t.nrep) { a = comb(a, v[r], mode); c = comb(c, w[r], mode); }
                }
                arr[i] = a;
                if (two) arr[i2] = c;
            }
        };
        if (need & NEED_FIRST) fold(t.first, 2);
        for (int c = 0; c < cv.ncols(); ++c) {
            if (need & NEED_SUM) fold(reinterpret_cast<unsigned long long*>(t.sum[c]), cv.is_f64(c) ? 1 : 0);
            if (need & NEED_CNT) fold(t.cnt[c], 0);
            if (need & NEED_MIN) fold(t.mn[c], 2);
            if (need & NEED_MAX) fold(t.mx[c], 3);
        }
    }
    __syncthreads();       // each thread reads back only what it wrote itself (same i striding) or after this barrier
    if (!fin.write_out) {
        // var / std: the host continues with the second pass; it needs the group count
        unsigned cntp = 0;
        for (unsigned i = tid; i < nslots; i += kSmThreads) cntp += __ldcg(&t.keys[i]) != kEmptyKey();
        atomicAdd(&s_counter, cntp);
        __syncthreads();
        if (tid == 0) fin.meta[0] = s_counter;
        publish();
        return;
    }
    if (drange && !fin.by_first) {
        // direct slots are already in key order: ordered compaction by a block scan over contiguous chunks
        const unsigned chunk = (nslots + kSmThreads - 1) / kSmThreads;
        const unsigned lo = tid * chunk, hi = lo + chunk < nslots ? lo + chunk : nslots;
        unsigned c = 0;
        for (unsigned i = lo; i < hi; ++i) c += __ldcg(&t.keys[i]) != kEmptyKey();
        unsigned incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        unsigned* wsum = reinterpret_cast<unsigned*>(s_red);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned basec = 0, tot = 0;
        for (int w = 0; w < kSmWarps; ++w) {
            const unsigned x = wsum[w];
            if (w < warp) basec += x;
            tot += x;
        }
        unsigned j = basec + incl - c;
        for (unsigned i = lo; i < hi; ++i) {
            if (__ldcg(&t.keys[i]) == kEmptyKey()) continue;
            write_group(t, cv.ncols(), fin.is_f64_mask, fin.agg_mask, fin.ddof, i, j, fin.ngcap, fin.out_keys, fin.out_vals);
            ++j;
        }
        if (tid == 0) fin.meta[0] = tot;
        publish();
        return;
    }
    // hashed slots (or sort=False): compact, bitonic sort of (key | first row, slot) in shared memory, write
    const unsigned sort_cap = (unsigned)(lay.alloc_bytes / 12);
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(sm);          // [sort_cap]
    unsigned* sslot = reinterpret_cast<unsigned*>(skey + sort_cap);                // [sort_cap]
    for (unsigned i = tid; i < nslots; i += kSmThreads) {
        const unsigned long long k = __ldcg(&t.keys[i]);
        if (k == kEmptyKey()) continue;
        const unsigned j = atomicAdd(&s_counter, 1u);
        if (j >= sort_cap) continue;
        unsigned long long sk;
        if (fin.by_first) sk = t.first[i];
        else {
            const unsigned long long key = (i == (unsigned)t.cap) ? kEmptyKey() : k;
            sk = fin.key_f64 ? f64_to_ordered(__longlong_as_double((long long)key)) : i64_to_ordered((long long)key);
        }
        skey[j] = sk;
        sslot[j] = i;
    }
    __syncthreads();
    const unsigned ng = s_counter < sort_cap ? s_counter : sort_cap;
    unsigned P = 1;
    while (P < ng) P <<= 1;
    for (unsigned i = ng + tid; i < P; i += kSmThreads) { skey[i] = ~0ull; sslot[i] = 0xffffffffu; }
    __syncthreads();
    for (unsigned k = 2; k <= P; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = tid; i < P; i += kSmThreads) {
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
    for (unsigned j = tid; j < ng; j += kSmThreads)
        write_group(t, cv.ncols(), fin.is_f64_mask, fin.agg_mask, fin.ddof, sslot[j], j, fin.ngcap, fin.out_keys, fin.out_vals);
    if (tid == 0) fin.meta[0] = ng;
    publish();
}

// second pass for var / std: ssd[slot] += (x - mean)^2  (nanvar's two passes, numpy_like.py:850-872)
template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_ssd_kernel(const void* __restrict__ keys, int64_t n, Cols cols, Table t,
                                                            unsigned long long dlo, unsigned long long drange) {
    const int tid = threadIdx.x;
    const int64_t ntiles = (n + kGbTile - 1) / kGbTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * kGbTile;
        unsigned long long kb[kRows], v0[kRows], g[kRows];
        __syncwarp();
        load4(keys, base, n, tid, cols.vec_ok, kb);
#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            unsigned long long k;
            g[e] = ~0ull;
            if (tile_row(base, tid, e) < n && canon_key<KEY_F64>(kb[e], k)) g[e] = global_find(t, k, dlo, drange);
            __syncwarp();
        }
        for (int c = 0; c < cols.ncols; ++c) {
            __syncwarp();
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

// ------------------------------------------------------------------ very high cardinality: sort, then segments
// When the table of a hash aggregate would be far larger than L2 (tens of millions of groups) every row costs
// two random HBM sectors; sorting (key, row) pairs with the onesweep radix sort and reducing equal-key runs
// streams instead, and leaves the groups in key order for free.  Rows are visited in sorted order:
//   gb_seg_count_kernel   per tile of 2048 sorted rows: number of run heads
//   (scan of the tile counts -> group id of every head, number of groups)
//   gb_seg_kernel         per thread 8 consecutive sorted rows: values gathered by row number, equal-key
//                         runs folded in registers, one atomic per run and aggregate into the group arrays
//                         (runs crossing thread or tile borders meet in the atomics)
// NaN keys sort last (sort.cu canon_key) and are skipped like in the hash path.  The stable sort keeps rows of a
// group in input order, so the head of a run is the group's first row (sort=False ordering).
// thresholds of the sort path (overridable for tests: SDCB200_GB_SORT_ROWS / SDCB200_GB_SORT_GROUPS)
int64_t sort_path_rows() {          // inputs below this keep the worst-case hash table
    const char* e = getenv("SDCB200_GB_SORT_ROWS");
    return e ? atoll(e) : (8ll << 20);
}
int64_t sort_path_groups() {        // expected groups above this go straight to the sort path
    const char* e = getenv("SDCB200_GB_SORT_GROUPS");
    return e ? atoll(e) : (4ll << 20);
}
int64_t dense_try_groups() {        // ... unless int64 keys turn out dense: direct table up to this many expected groups
    const char* e = getenv("SDCB200_GB_DENSE_GROUPS");
    return e ? atoll(e) : (256ll << 20);
}
constexpr int kSegItems = 8;
constexpr int kSegTile = kGbThreads * kSegItems;

template <bool KEY_F64>
__device__ __forceinline__ bool seg_key(unsigned long long raw, bool skip_neg, unsigned long long& k) {
    return canon_key<KEY_F64>(raw, k) && !(skip_neg && (long long)k < 0);
}

template <bool KEY_F64>
__device__ __forceinline__ unsigned seg_heads(const unsigned long long* __restrict__ sk, int64_t i0, int64_t n, bool skip_neg,
                                              unsigned long long k[kSegItems], unsigned& valid) {
    // bit e of the result: row i0 + e starts a run; bit e of valid: row exists and its key is not NaN
    unsigned heads = 0;
    valid = 0;
    unsigned long long prev = 0;
    bool have_prev = false;
    if (i0 > 0 && i0 < n) have_prev = seg_key<KEY_F64>(sk[i0 - 1], skip_neg, prev);
#pragma unroll
    for (int e = 0; e < kSegItems; ++e) {
        k[e] = 0;
        if (i0 + e < n && seg_key<KEY_F64>(sk[i0 + e], skip_neg, k[e])) {
            valid |= 1u << e;
            if (!have_prev || k[e] != prev) heads |= 1u << e;
            prev = k[e];
            have_prev = true;
        }
    }
    return heads;
}

template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_seg_count_kernel(const unsigned long long* __restrict__ sk, int64_t n,
                                                                  int skip_neg, unsigned long long* __restrict__ tile_counts) {
    __shared__ unsigned wcnt[kGbThreads / 32];
    const int64_t i0 = (int64_t)blockIdx.x * kSegTile + (int64_t)threadIdx.x * kSegItems;
    unsigned long long k[kSegItems];
    unsigned valid;
    unsigned c = __popc(seg_heads<KEY_F64>(sk, i0, n, skip_neg != 0, k, valid));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < kGbThreads / 32; ++w) t += wcnt[w];
        tile_counts[blockIdx.x] = t;
    }
}

// pass 0: sum / count / min / max / first into the group arrays; pass 1: sum of squared deviations (var / std)
// VALPAY: the sort carried the (single) value column itself as its payload, so `rows` holds the values in sorted
// order and nothing is gathered -- a random 8-byte gather costs a whole 128-byte DRAM burst per row.
template <bool KEY_F64, typename ROW, bool VALPAY>
__global__ void __launch_bounds__(kGbThreads) gb_seg_kernel(const unsigned long long* __restrict__ sk,
                                                            const ROW* __restrict__ rows, int64_t n, Cols cols, Table t,
                                                            int need, int pass, int skip_neg,
                                                            const unsigned long long* __restrict__ tile_offs) {
    __shared__ unsigned long long wsum[kGbThreads / 32];
    const int64_t i0 = (int64_t)blockIdx.x * kSegTile + (int64_t)threadIdx.x * kSegItems;
    unsigned long long k[kSegItems];
    unsigned valid;
    const unsigned heads = seg_heads<KEY_F64>(sk, i0, n, skip_neg != 0, k, valid);
    unsigned long long tot;
    // group id of the run that is open when this thread starts (= heads before it, minus one)
    unsigned long long gid = tile_offs[blockIdx.x] + block_excl_scan_u64(__popc(heads), &tot, wsum) - 1ull;
    if (!valid) return;
    ROW r[kSegItems];
#pragma unroll
    for (int e = 0; e < kSegItems; ++e) r[e] = (valid >> e) & 1u ? rows[i0 + e] : ROW(0);
    if (pass == 0) {
        unsigned long long g = gid;
#pragma unroll
        for (int e = 0; e < kSegItems; ++e) {
            if ((heads >> e) & 1u) {
                ++g;
                t.keys[g] = k[e];
                if (t.start) t.start[g] = (unsigned long long)(i0 + e);
                if (!VALPAY && (need & NEED_FIRST)) t.first[g] = (unsigned long long)r[e];
            }
        }
    }
    for (int c = 0; c < cols.ncols; ++c) {
        const bool f = cols.is_f64[c] != 0;
        const unsigned long long* col = reinterpret_cast<const unsigned long long*>(cols.data[c]);
        unsigned long long bits[kSegItems];
#pragma unroll
        for (int e = 0; e < kSegItems; ++e)          // gathers in flight together
            bits[e] = (valid >> e) & 1u ? (VALPAY ? (unsigned long long)r[e] : __ldg(col + r[e])) : 0ull;
        unsigned long long g = gid;
        // open run: folded in registers, flushed with atomics when the next head (or the thread's end) comes
        double fs = 0.0, fp = 1.0;
        long long is = 0, ip = 1;
        unsigned long long cn = 0, mn = ~0ull, mx = 0ull;
        bool open = false;
        auto flush = [&]() {
            if (!open) return;
            if (pass == 1) { if (cn) atomicAdd(&t.ssd[c][g], fs); }
            else {
                if ((need & NEED_PROD) && cn) {
                    // no native atomic multiply: CAS loop (runs rarely meet: only at thread / tile borders)
                    unsigned long long* p = reinterpret_cast<unsigned long long*>(t.prod[c]) + g;
                    unsigned long long old = *p, assumed;
                    do {
                        assumed = old;
                        const unsigned long long nv = f ? (unsigned long long)__double_as_longlong(__longlong_as_double((long long)assumed) * fp)
                                                        : (unsigned long long)((long long)assumed * ip);
                        old = atomicCAS(p, assumed, nv);
                    } while (old != assumed);
                }
                if (need & NEED_SUM) {
                    if (f) { if (cn) atomicAdd(reinterpret_cast<double*>(t.sum[c]) + g, fs); }
                    else atomicAdd(reinterpret_cast<unsigned long long*>(t.sum[c]) + g, (unsigned long long)is);
                }
                if ((need & NEED_CNT) && cn) atomicAdd(&t.cnt[c][g], cn);
                if ((need & NEED_MIN) && cn) atomicMin(&t.mn[c][g], mn);
                if ((need & NEED_MAX) && cn) atomicMax(&t.mx[c][g], mx);
            }
            fs = 0.0; is = 0; cn = 0; mn = ~0ull; mx = 0ull; fp = 1.0; ip = 1;
        };
#pragma unroll
        for (int e = 0; e < kSegItems; ++e) {
            if (!((valid >> e) & 1u)) continue;
            if ((heads >> e) & 1u) { flush(); ++g; }
            open = true;
            if (pass == 1) {
                const double cnt = (double)t.cnt[c][g];
                double v, mean;
                if (f) {
                    v = __longlong_as_double((long long)bits[e]);
                    if (v != v) continue;
                    mean = reinterpret_cast<const double*>(t.sum[c])[g] / cnt;
                } else {
                    v = (double)(long long)bits[e];
                    mean = (double)reinterpret_cast<const long long*>(t.sum[c])[g] / cnt;
                }
                const double d = v - mean;
                fs += d * d;
                ++cn;
            } else if (f) {
                const double v = __longlong_as_double((long long)bits[e]);
                if (v != v) continue;                                    // skipna
                fs += v;
                fp *= v;
                ++cn;
                const unsigned long long o = f64_to_ordered(v);
                mn = o < mn ? o : mn;
                mx = o > mx ? o : mx;
            } else {
                is += (long long)bits[e];
                ip = (long long)((unsigned long long)ip * bits[e]);      // wraps like numpy's int64 product
                ++cn;
                const unsigned long long o = i64_to_ordered((long long)bits[e]);
                mn = o < mn ? o : mn;
                mx = o > mx ? o : mx;
            }
        }
        flush();
    }
}

// exact min / max of an int64 key column (order-preserving unsigned image)
__global__ void __launch_bounds__(256) gb_key_minmax_kernel(const long long* __restrict__ keys, int64_t n,
                                                            unsigned long long* __restrict__ mm /* [0] min (init ~0), [1] max (init 0) */) {
    unsigned long long mn = ~0ull, mx = 0ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long u = (unsigned long long)keys[i] ^ 0x8000000000000000ull;
        mn = u < mn ? u : mn;
        mx = u > mx ? u : mx;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, d), b = __shfl_xor_sync(0xffffffffu, mx, d);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm, mn);
        atomicMax(mm + 1, mx);
    }
}

// every table array gets its initial value in one launch
struct InitPlan {
    unsigned long long* ptr[4 + 5 * kMaxCols];
    unsigned long long val[4 + 5 * kMaxCols];
    unsigned long long len[4 + 5 * kMaxCols];
    int count;
};
__global__ void gb_init_kernel(InitPlan plan) {
    for (int a = blockIdx.y; a < plan.count; a += gridDim.y) {
        unsigned long long* p = plan.ptr[a];
        const unsigned long long v = plan.val[a], n = plan.len[a];
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
             i += (unsigned long long)gridDim.x * blockDim.x)
            p[i] = v;
    }
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

// ---- median (sort-and-segment path): rows ordered by (key, value), NaN values last inside every group ----
__global__ void gb_gather_u32_kernel(const unsigned long long* __restrict__ src, const uint32_t* __restrict__ idx,
                                     unsigned long long* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}

// numpy.nanmedian per group (Series.median, sdc/datatypes/hpat_pandas_series_functions.py:3721-3730): the mean of the
// two middle non-NaN values (the same value twice for an odd count); NaN for a group without one
__global__ void gb_median_kernel(Table t, int c, int is_f64, const unsigned long long* __restrict__ col,
                                 const uint32_t* __restrict__ rows_kv, int64_t ng) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < ng; g += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long m = t.cnt[c][g];
        double r = nan;
        if (m) {
            const unsigned long long s0 = t.start[g];
            const unsigned long long a = col[rows_kv[s0 + (m - 1) / 2]], b = col[rows_kv[s0 + m / 2]];
            const double x = is_f64 ? __longlong_as_double((long long)a) : (double)(long long)a;
            const double y = is_f64 ? __longlong_as_double((long long)b) : (double)(long long)b;
            r = (x + y) / 2.0;
        }
        t.med[c][g] = r;
    }
}

// segment path: the group arrays are already dense (group j at index j, in key order); `order` (sort=False)
// permutes them into first-appearance order
__global__ void gb_finalize_seg_kernel(Table t, int ncols, unsigned is_f64_mask, unsigned agg_mask, long long ddof,
                                       const int64_t* __restrict__ order, int64_t ng, unsigned long long* out_keys,
                                       unsigned long long* out_vals) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ng; j += (int64_t)gridDim.x * blockDim.x)
        write_group(t, ncols, is_f64_mask, agg_mask, ddof, (unsigned long long)(order ? order[j] : j), j, ng, out_keys, out_vals);
}

// ------------------------------------------------------------------ host side
// composite group ids of several key columns: out = a * nb + b, `invalid` where either id is negative
__global__ void combine_ids_kernel(const long long* __restrict__ a, const long long* __restrict__ b, long long nb, long long invalid,
                                   long long* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long x = a[i], y = b[i];
        out[i] = (x < 0 || y < 0) ? invalid : x * nb + y;
    }
}

struct PendingSmall;
struct GroupbyResult {
    int key_dtype = 0;
    int ncols = 0;
    int64_t ngroups = 0;
    int64_t ngcap = 0;                 // row stride of `vals`
    unsigned agg_mask = 0;
    unsigned long long* keys = nullptr;    // [ngcap]
    unsigned long long* vals = nullptr;    // [popc(agg_mask) * ncols * ngcap]
    bool one_block = false;                // vals lives in the allocation keys points at (low-cardinality path)
    bool lean_direct = false;              // produced by one of gb_smem_kernel's lean loops (direct window or hashed slots)
    struct PendingSmall* pending = nullptr;   // sdcb200_groupby_begin: the kernel is in flight, sdcb200_groupby_end finishes
    cudaStream_t stream = nullptr;            // the caller's stream the result was produced on: the buffers are handed out
                                              // as zero-copy views and consumed there, so they are freed in ITS order
                                              // (the host-buffer forms run on streams of their own, synchronise them and
                                              // reset this to the default stream before they destroy them)
};

std::mutex g_mu;
std::unordered_map<int64_t, GroupbyResult*> g_results;
int64_t g_next_handle = 1;

void free_pending(GroupbyResult* r);
void free_result(GroupbyResult* r) {
    if (!r) return;
    free_pending(r);
    if (r->keys) dev_free(r->keys, r->stream);
    if (r->vals && !r->one_block) dev_free(r->vals, r->stream);
    delete r;
}

// SDCB200_GB_LEAN=0 keeps the generic loop of gb_smem_kernel for dense keys too (A/B measurements)
bool gb_lean_enabled() {
    static const int on = [] {
        const char* e = getenv("SDCB200_GB_LEAN");
        return (e && *e) ? atoi(e) : 1;
    }();
    return on != 0;
}

// tuning knobs of the lean loop, read on every call so that one process can compare settings
int gb_env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

int need_for(unsigned agg_mask, bool sort) {
    int need = 0;
    if (agg_mask & (AGG_SUM | AGG_MEAN | AGG_VAR | AGG_STD)) need |= NEED_SUM;
    if (agg_mask & (AGG_MEAN | AGG_MIN | AGG_MAX | AGG_COUNT | AGG_VAR | AGG_STD | AGG_MEDIAN)) need |= NEED_CNT;
    if (agg_mask & AGG_PROD) need |= NEED_PROD;
    if (agg_mask & AGG_MIN) need |= NEED_MIN;
    if (agg_mask & AGG_MAX) need |= NEED_MAX;
    if (!sort) need |= NEED_FIRST;
    return need;
}

SmemLayout make_layout(int ncols, int need, int want_slots, int budget_bytes, unsigned long long cap) {
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
    // the last CTA's bitonic sort holds up to `cap` (key, slot) pairs of 12 bytes in the same allocation
    const int sort_bytes = (int)cap * 12;
    L.alloc_bytes = off > sort_bytes ? off : sort_bytes;
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
    void* detach() {                       // the caller takes over the allocation (dev_free on the same stream later)
        void* p = block.p;
        block.p = nullptr;
        return p;
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

// pinned landing zone for the control-block read-back (a pageable destination makes the copy a staged one)
unsigned long long* pinned_meta() {
    static thread_local unsigned long long* p = nullptr;
    if (!p && cudaHostAlloc(reinterpret_cast<void**>(&p), 256, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) p = nullptr;
    return p;
}

// the address a kernel stores to so that pinned_meta() sees it (null when there is no mapped buffer)
unsigned long long* pinned_meta_device() {
    unsigned long long* h = pinned_meta();
    void* d = nullptr;
    if (!h || cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return reinterpret_cast<unsigned long long*>(d);
}

// ---- asynchronous low-cardinality groupby (sdcb200_groupby_begin / _end) ----
// A call in flight owns a 256-byte slot of mapped pinned memory (the control block the kernel's last CTA publishes) and
// the table it aggregates into; sdcb200_groupby_end waits for the stream, reads the slot and finishes the result.
struct PinnedSlots {
    std::mutex mu;
    std::vector<unsigned long long*> free_host;          // host addresses of free 256-byte slots
    unsigned long long* take() {
        std::lock_guard<std::mutex> lock(mu);
        if (free_host.empty()) {
            unsigned long long* chunk = nullptr;
            if (cudaHostAlloc(reinterpret_cast<void**>(&chunk), 64 * 256, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            for (int i = 0; i < 64; ++i) free_host.push_back(chunk + i * 32);
        }
        unsigned long long* p = free_host.back();
        free_host.pop_back();
        return p;
    }
    void give(unsigned long long* p) {
        if (!p) return;
        std::lock_guard<std::mutex> lock(mu);
        free_host.push_back(p);
    }
};
PinnedSlots g_slots;

struct PendingSmall {
    unsigned long long* slot = nullptr;     // pinned control-block copy (host address)
    void* table = nullptr;                  // the table allocation (stream-ordered), released by _end
    cudaStream_t stream = nullptr;
    bool lean = false, lean_hash = false;
    // what a rerun needs when the optimistic attempt overflowed (a key outside the sampled window, too many groups):
    // the caller keeps the input columns alive until sdcb200_groupby_end
    const void* keys = nullptr;
    int64_t n = 0;
    Cols cols;
    unsigned agg_mask = 0;
    bool sort = true;
    long long ddof = 1;
    int key_dtype = 0, key_mode = 0;
    int64_t expected_groups = 0;
};

void free_pending(GroupbyResult* r) {        // a handle dropped without sdcb200_groupby_end: wait, then release
    PendingSmall* ps = r->pending;
    if (!ps) return;
    r->pending = nullptr;
    cudaStreamSynchronize(ps->stream);
    g_slots.give(ps->slot);
    dev_free(ps->table, ps->stream);
    delete ps;
}

unsigned long long pow2_at_least(unsigned long long v, unsigned long long lo) {
    unsigned long long c = lo;
    while (c < v) c <<= 1;
    return c;
}

// very high cardinality: radix sort of (key, row), then one streaming pass over equal-key runs
template <bool KEY_F64>
int32_t run_groupby_sorted(const void* keys, int64_t n, Cols cols, unsigned agg_mask, bool sort, long long ddof,
                           int key_dtype, int key_mode, cudaStream_t s, GroupbyResult* res) {
    const int skip_neg = key_mode == KEYS_GID ? 1 : 0;
    const int need = need_for(agg_mask, sort);
    const bool need_ssd = (agg_mask & (AGG_VAR | AGG_STD)) != 0;
    const bool need_med = (agg_mask & AGG_MEDIAN) != 0;
    if (need_med && n > 0xffffffffll) { set_error("groupby: median needs fewer than 2^32 rows"); return SDCB200_ERR_ARG; }
    const int sms = sm_count();
    const int nagg = __builtin_popcount(agg_mask);
    unsigned is_f64_mask = 0;
    for (int c = 0; c < cols.ncols; ++c) if (cols.is_f64[c]) is_f64_mask |= 1u << c;
    const bool small_rows = n <= 0xffffffffll;
    // one value column and no first-row bookkeeping: the values ride through the sort as the payload
    const bool valpay = cols.ncols == 1 && !(need & NEED_FIRST);
    const int pay = valpay ? 8 : (small_rows ? 4 : 8);
    Scratch ka, kb, pa, pb, tcnt;
    SDC_TRY(ka.alloc((size_t)n * 8, s));
    SDC_TRY(kb.alloc((size_t)n * 8, s));
    SDC_TRY(pa.alloc((size_t)n * pay, s));
    SDC_TRY(pb.alloc((size_t)n * pay, s));
    const void *rk, *rp;
    SDC_TRY(radix_sort_core(key_dtype, keys, ka.p, kb.p, pay, valpay ? cols.data[0] : nullptr, pa.p, pb.p, n, !valpay, false,
                            &rk, &rp, s));
    if (valpay && rp == nullptr) rp = cols.data[0];          // no pass ran (every digit constant): the column as it is
    const unsigned long long* sk = reinterpret_cast<const unsigned long long*>(rk);
    const int64_t nt = ceil_div(n, kSegTile);
    SDC_TRY(tcnt.alloc((size_t)(nt + 1) * 8, s));
    gb_seg_count_kernel<KEY_F64><<<(unsigned)nt, kGbThreads, 0, s>>>(sk, n, skip_neg, tcnt.as<unsigned long long>());
    SDC_CHECK_LAUNCH();
    scan_tiles_kernel<<<1, kScanTilesThreads, 0, s>>>(tcnt.as<unsigned long long>(), nt);
    SDC_CHECK_LAUNCH();
    unsigned long long total = 0;
    SDC_CUDA_TRY(cudaMemcpyAsync(&total, tcnt.as<unsigned long long>() + nt, 8, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    const int64_t ng = (int64_t)total;
    res->ngroups = ng;
    res->ngcap = ng;
    if (ng == 0) return SDCB200_OK;

    TableMem mem;
    Table t;
    memset(&t, 0, sizeof(t));
    t.cap = ~0ull;                    // no spare slot: group j lives at index j
    t.nrep = 1;
    InitPlan ip;
    ip.count = 0;
    {
        int narr = 1 + ((need & NEED_FIRST) ? 1 : 0) + (need_med ? 1 : 0);
        narr += cols.ncols * (((need & NEED_SUM) ? 1 : 0) + ((need & NEED_CNT) ? 1 : 0) + ((need & NEED_MIN) ? 1 : 0) +
                              ((need & NEED_MAX) ? 1 : 0) + (need_ssd ? 1 : 0) + ((need & NEED_PROD) ? 1 : 0) + (need_med ? 1 : 0));
        SDC_TRY(mem.reserve((size_t)narr * (((size_t)ng * 8 + 255) & ~size_t(255)) + 256, s));
    }
    void* p;
    auto add = [&](void** dst, unsigned long long init, bool do_init) -> int32_t {
        SDC_TRY(mem.get(&p, (size_t)ng * 8, s));
        *dst = p;
        if (do_init) {
            ip.ptr[ip.count] = reinterpret_cast<unsigned long long*>(p);
            ip.len[ip.count] = (unsigned long long)ng;
            ip.val[ip.count++] = init;
        }
        return SDCB200_OK;
    };
    SDC_TRY(add(reinterpret_cast<void**>(&t.keys), 0ull, false));                 // every entry is written by its head
    if (need & NEED_FIRST) SDC_TRY(add(reinterpret_cast<void**>(&t.first), 0ull, false));
    if (need_med) SDC_TRY(add(reinterpret_cast<void**>(&t.start), 0ull, false));
    for (int c = 0; c < cols.ncols; ++c) {
        if (need & NEED_SUM) SDC_TRY(add(&t.sum[c], 0ull, true));
        if (need & NEED_CNT) SDC_TRY(add(reinterpret_cast<void**>(&t.cnt[c]), 0ull, true));
        if (need & NEED_MIN) SDC_TRY(add(reinterpret_cast<void**>(&t.mn[c]), ~0ull, true));
        if (need & NEED_MAX) SDC_TRY(add(reinterpret_cast<void**>(&t.mx[c]), 0ull, true));
        if (need_ssd) SDC_TRY(add(reinterpret_cast<void**>(&t.ssd[c]), 0ull, true));
        if (need & NEED_PROD) SDC_TRY(add(&t.prod[c], cols.is_f64[c] ? 0x3ff0000000000000ull : 1ull, true));   // 1.0 / 1
        if (need_med) SDC_TRY(add(reinterpret_cast<void**>(&t.med[c]), 0ull, false));
    }
    if (ip.count) {
        unsigned gx = (unsigned)(ceil_div(ng, 1024) < (int64_t)sms * 4 ? ceil_div(ng, 1024) : (int64_t)sms * 4);
        dim3 grid(gx ? gx : 1, ip.count < 16 ? ip.count : 16);
        gb_init_kernel<<<grid, 256, 0, s>>>(ip);
        SDC_CHECK_LAUNCH();
    }
    for (int pass = 0; pass < (need_ssd ? 2 : 1); ++pass) {
        if (valpay)
            gb_seg_kernel<KEY_F64, unsigned long long, true><<<(unsigned)nt, kGbThreads, 0, s>>>(
                sk, reinterpret_cast<const unsigned long long*>(rp), n, cols, t, need, pass, skip_neg, tcnt.as<unsigned long long>());
        else if (small_rows)
            gb_seg_kernel<KEY_F64, uint32_t, false><<<(unsigned)nt, kGbThreads, 0, s>>>(sk, reinterpret_cast<const uint32_t*>(rp), n, cols, t,
                                                                                       need, pass, skip_neg, tcnt.as<unsigned long long>());
        else
            gb_seg_kernel<KEY_F64, unsigned long long, false><<<(unsigned)nt, kGbThreads, 0, s>>>(
                sk, reinterpret_cast<const unsigned long long*>(rp), n, cols, t, need, pass, skip_neg, tcnt.as<unsigned long long>());
        SDC_CHECK_LAUNCH();
    }
    if (need_med) {
        // rows ordered by (key, value): stable sort by value, then stable sort of the permuted keys carrying the rows.
        // The key order is the same as in the main sort, so its group starts and counts apply unchanged.
        Scratch va, vb, r1a, r1b, kp, k2a, k2b, r2a, r2b;
        SDC_TRY(va.alloc((size_t)n * 8, s));
        SDC_TRY(vb.alloc((size_t)n * 8, s));
        SDC_TRY(r1a.alloc((size_t)n * 4, s));
        SDC_TRY(r1b.alloc((size_t)n * 4, s));
        SDC_TRY(kp.alloc((size_t)n * 8, s));
        SDC_TRY(k2a.alloc((size_t)n * 8, s));
        SDC_TRY(k2b.alloc((size_t)n * 8, s));
        SDC_TRY(r2a.alloc((size_t)n * 4, s));
        SDC_TRY(r2b.alloc((size_t)n * 4, s));
        const int ggrid = (int)(ceil_div(n, 256) < (int64_t)sms * 8 ? ceil_div(n, 256) : (int64_t)sms * 8);
        const int mgrid = (int)(ceil_div(ng, 256) < (int64_t)sms * 8 ? ceil_div(ng, 256) : (int64_t)sms * 8);
        for (int c = 0; c < cols.ncols; ++c) {
            const void *vk, *vr, *k2, *r2;
            SDC_TRY(radix_sort_core(cols.is_f64[c] ? SDCB200_F64 : SDCB200_I64, cols.data[c], va.p, vb.p, 4, nullptr, r1a.p, r1b.p, n,
                                    true, false, &vk, &vr, s));
            gb_gather_u32_kernel<<<ggrid, 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(keys),
                                                       reinterpret_cast<const uint32_t*>(vr), kp.as<unsigned long long>(), n);
            SDC_CHECK_LAUNCH();
            SDC_TRY(radix_sort_core(key_dtype, kp.p, k2a.p, k2b.p, 4, vr, r2a.p, r2b.p, n, false, false, &k2, &r2, s));
            gb_median_kernel<<<mgrid, 256, 0, s>>>(t, c, cols.is_f64[c], reinterpret_cast<const unsigned long long*>(cols.data[c]),
                                                   reinterpret_cast<const uint32_t*>(r2), ng);
            SDC_CHECK_LAUNCH();
        }
        SDC_CUDA_TRY(cudaStreamSynchronize(s));       // the scratch above goes out of scope
    }
    Scratch order;
    if (!sort) {
        SDC_TRY(order.alloc((size_t)ng * 8, s));
        SDC_TRY(argsort_device(SDCB200_U64, t.first, ng, true, order.as<int64_t>(), s));
    }
    SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->keys), (size_t)ng * 8, s));
    SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->vals), (size_t)ng * 8 * (nagg * cols.ncols + 1), s));
    const int fgrid = (int)(ceil_div(ng, 256) < (int64_t)sms * 8 ? ceil_div(ng, 256) : (int64_t)sms * 8);
    gb_finalize_seg_kernel<<<fgrid, 256, 0, s>>>(t, cols.ncols, is_f64_mask, agg_mask, ddof,
                                                 sort ? nullptr : order.as<int64_t>(), ng, res->keys, res->vals);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

// ------------------------------------------------------------------ dense window path
// int64 keys that fill a DENSE range [lo, lo + range) with tens of millions of groups, one float64 column, aggregates out
// of {sum, mean, min, max, count}, sort=True.  The table is direct-addressed (slot = key - lo) and holds only what the
// aggregates need -- no key words: a slot's key is lo + slot, and a slot is occupied when its count is non-zero (without a
// count array: when its sum left the initial -0.0), or when a row that changes neither (a NaN value; -0.0 for sum alone)
// stored its mark byte.  So a row costs ONE fire-and-forget L2 reduction per aggregate and nothing else.
//   gb_key_minmax_kernel   exact key range (one read of the keys)
//   range_partition_pass   (tables beyond L2 only) ONE unordered onesweep pass of (key, value) into <= 256 key-range
//                          buckets: the rows then walk the table slice by slice and the slice in use stays in L2
//   gb_window_kernel       the reductions; CTAs take tiles in row order
//   gb_window_count_kernel / scan_tiles_kernel / gb_window_write_kernel
//                          occupied slots counted per tile of 2048 slots, scanned, written in slot = key order
struct WinTable {
    double* sum;                    // [range] init -0.0 (sum alone) / 0.0
    unsigned* cnt;                  // [range] non-NaN rows
    unsigned long long* mn;         // [range] ordered u64, init ~0
    unsigned long long* mx;         // [range] ordered u64, init 0
    unsigned char* mark;            // [range] 1 = a row of this key left the other arrays untouched
    unsigned long long lo, range;
};
constexpr unsigned long long kWinNegZero = 0x8000000000000000ull;

__global__ void __launch_bounds__(kGbThreads) gb_window_kernel(const unsigned long long* __restrict__ keys,
                                                              const unsigned long long* __restrict__ vals, int64_t n, int vec_ok,
                                                              WinTable w, int need, unsigned long long* __restrict__ ticket) {
    // Tiles are handed out by a ticket counter, not by a fixed stride: CTAs drift apart over a hundred tiles each, and
    // with a stride the rows in flight then span dozens of key-range buckets -- tens of megabytes of table -- instead of
    // the one or two that the resident CTAs cover when they always take the next tile in row order.
    __shared__ unsigned long long s_tile[2];
    const int tid = threadIdx.x;
    const int64_t ntiles = (n + kGbTile - 1) / kGbTile;
    if (tid == 0) s_tile[0] = atomicAdd(ticket, 1ull);
    __syncthreads();
    for (int it = 0;; ++it) {
        const int64_t tile = (int64_t)s_tile[it & 1];
        if (tile >= ntiles) break;
        const int64_t base = tile * kGbTile;
        unsigned long long kb[kRows], vb[kRows];
        load4(keys, base, n, tid, vec_ok, kb);
        load4(vals, base, n, tid, vec_ok, vb);
        if (tid == 0) s_tile[(it + 1) & 1] = atomicAdd(ticket, 1ull);      // the next tile, read after the barrier below
#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            if (tile_row(base, tid, e) >= n) continue;
            const unsigned long long d = kb[e] - w.lo;
            if (d >= w.range) continue;                              // cannot happen: the range is exact
            const double v = __longlong_as_double((long long)vb[e]);
            if (v != v) { w.mark[d] = 1; continue; }                 // skipna: the group exists, nothing is added
            if (need & NEED_SUM) atomicAdd(w.sum + d, v);
            if (need & NEED_CNT) atomicAdd(w.cnt + d, 1u);
            else if (vb[e] == kWinNegZero) w.mark[d] = 1;            // sum alone: -0.0 + -0.0 stays -0.0
            if (need & NEED_MIN) atomicMin(w.mn + d, f64_to_ordered(v));
            if (need & NEED_MAX) atomicMax(w.mx + d, f64_to_ordered(v));
        }
        __syncthreads();
    }
}

__device__ __forceinline__ bool win_present(const WinTable& w, int need, unsigned long long d) {
    if (w.mark[d]) return true;
    if (need & NEED_CNT) return w.cnt[d] != 0u;
    return reinterpret_cast<const unsigned long long*>(w.sum)[d] != kWinNegZero;
}

constexpr int kWinTile = 2048;          // slots per CTA of the count / write kernels: 256 threads x 8 consecutive slots
__global__ void __launch_bounds__(256) gb_window_count_kernel(WinTable w, int need, unsigned long long* __restrict__ tcnt) {
    __shared__ unsigned s_w[8];
    const unsigned long long d0 = (unsigned long long)blockIdx.x * kWinTile + (unsigned long long)threadIdx.x * 8;
    unsigned c = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (d0 + e < w.range && win_present(w, need, d0 + e)) ++c;
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) c += __shfl_xor_sync(0xffffffffu, c, x);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += s_w[i];
        tcnt[blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(256) gb_window_write_kernel(WinTable w, int need, unsigned agg_mask,
                                                             const unsigned long long* __restrict__ tcnt, int64_t ng,
                                                             unsigned long long* __restrict__ out_keys,
                                                             unsigned long long* __restrict__ out_vals) {
    __shared__ unsigned s_w[8];
    const unsigned long long d0 = (unsigned long long)blockIdx.x * kWinTile + (unsigned long long)threadIdx.x * 8;
    unsigned pres = 0, c = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (d0 + e < w.range && win_present(w, need, d0 + e)) { pres |= 1u << e; ++c; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = c;
#pragma unroll
    for (int x = 1; x < 32; x <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, x);
        if (lane >= x) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < warp) wbase += s_w[i];
    int64_t j = (int64_t)tcnt[blockIdx.x] + wbase + incl - c;
    const unsigned long long nanb = 0x7ff8000000000000ull;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (!((pres >> e) & 1u)) continue;
        const unsigned long long d = d0 + e;
        out_keys[j] = w.lo + d;
        const unsigned cn = (need & NEED_CNT) ? w.cnt[d] : 0u;
        const double sm = (need & NEED_SUM) ? w.sum[d] + 0.0 : 0.0;       // -0.0 (nothing added) reads as the reference's 0.0
        int rank = 0;
        if (agg_mask & AGG_SUM) out_vals[(int64_t)rank++ * ng + j] = (unsigned long long)__double_as_longlong(sm);
        if (agg_mask & AGG_MEAN) out_vals[(int64_t)rank++ * ng + j] = cn ? (unsigned long long)__double_as_longlong(sm / (double)cn) : nanb;
        if (agg_mask & AGG_MIN) out_vals[(int64_t)rank++ * ng + j] = cn ? (unsigned long long)__double_as_longlong(ordered_to_f64(w.mn[d])) : nanb;
        if (agg_mask & AGG_MAX) out_vals[(int64_t)rank++ * ng + j] = cn ? (unsigned long long)__double_as_longlong(ordered_to_f64(w.mx[d])) : nanb;
        if (agg_mask & AGG_COUNT) out_vals[(int64_t)rank++ * ng + j] = (unsigned long long)cn;
        ++j;
    }
}

// *done = false (and SDCB200_OK): the keys are not dense, or the table does not fit -- the caller sorts instead
int32_t run_groupby_window(const void* keys, int64_t n, const Cols& cols, unsigned agg_mask, int64_t expected_groups,
                           cudaStream_t s, GroupbyResult* res, bool* done) {
    *done = false;
    const int need = need_for(agg_mask, true);
    const int sms = sm_count();
    const int nagg = __builtin_popcount(agg_mask);
    Scratch mm;
    SDC_TRY(mm.alloc(32, s));
    const unsigned long long init[4] = {~0ull, 0ull, 0ull, 0ull};           // min, max, the tile ticket of gb_window_kernel
    SDC_CUDA_TRY(cudaMemcpyAsync(mm.p, init, 32, cudaMemcpyHostToDevice, s));
    gb_key_minmax_kernel<<<sms * 8, 256, 0, s>>>(reinterpret_cast<const long long*>(keys), n, mm.as<unsigned long long>());
    SDC_CHECK_LAUNCH();
    unsigned long long h[2];
    SDC_CUDA_TRY(cudaMemcpyAsync(h, mm.p, 16, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    const unsigned long long span = h[1] - h[0];                     // ordered images: no overflow
    if (span >= 4ull * (unsigned long long)expected_groups || span >= (1ull << 30)) return SDCB200_OK;
    WinTable w;
    memset(&w, 0, sizeof(w));
    w.lo = h[0] ^ 0x8000000000000000ull;
    w.range = span + 1;
    const size_t r8 = ((size_t)w.range * 8 + 255) & ~size_t(255), r4 = ((size_t)w.range * 4 + 255) & ~size_t(255),
                 r1 = ((size_t)w.range + 255) & ~size_t(255);
    const size_t table_bytes = ((need & NEED_SUM) ? r8 : 0) + ((need & NEED_CNT) ? r4 : 0) + ((need & NEED_MIN) ? r8 : 0) +
                               ((need & NEED_MAX) ? r8 : 0) + r1;
    TableMem mem;
    if (mem.reserve(table_bytes + 256, s) != SDCB200_OK) { cudaGetLastError(); return SDCB200_OK; }
    InitPlan ip;
    ip.count = 0;
    void* p;
    if (need & NEED_SUM) {
        SDC_TRY(mem.get(&p, r8, s));
        w.sum = reinterpret_cast<double*>(p);
        ip.ptr[ip.count] = reinterpret_cast<unsigned long long*>(p);
        ip.len[ip.count] = w.range;
        ip.val[ip.count++] = (need & NEED_CNT) ? 0ull : kWinNegZero;
    }
    if (need & NEED_MIN) {
        SDC_TRY(mem.get(&p, r8, s));
        w.mn = reinterpret_cast<unsigned long long*>(p);
        ip.ptr[ip.count] = w.mn;
        ip.len[ip.count] = w.range;
        ip.val[ip.count++] = ~0ull;
    }
    if (need & NEED_MAX) {
        SDC_TRY(mem.get(&p, r8, s));
        w.mx = reinterpret_cast<unsigned long long*>(p);
        SDC_CUDA_TRY(cudaMemsetAsync(p, 0, r8, s));
    }
    if (need & NEED_CNT) {
        SDC_TRY(mem.get(&p, r4, s));
        w.cnt = reinterpret_cast<unsigned*>(p);
        SDC_CUDA_TRY(cudaMemsetAsync(p, 0, r4, s));
    }
    SDC_TRY(mem.get(&p, r1, s));
    w.mark = reinterpret_cast<unsigned char*>(p);
    SDC_CUDA_TRY(cudaMemsetAsync(p, 0, r1, s));
    if (ip.count) {
        unsigned gx = (unsigned)(ceil_div((int64_t)w.range, 1024) < (int64_t)sms * 4 ? ceil_div((int64_t)w.range, 1024) : (int64_t)sms * 4);
        gb_init_kernel<<<dim3(gx ? gx : 1, ip.count), 256, 0, s>>>(ip);
        SDC_CHECK_LAUNCH();
    }
    // a table around the size of L2 takes the rows as they come; a larger one gets them bucketed by key range first
    // (SDCB200_GB_WIN_PART: 0 / 1 force it off / on)
    const int part_env = gb_env_int("SDCB200_GB_WIN_PART", -1);
    const bool part = part_env >= 0 ? part_env != 0 : table_bytes > (size_t(48) << 20);
    Scratch part_keys, part_vals, part_begin;
    const void* k = keys;
    const void* v = cols.data[0];
    int vec_ok = cols.vec_ok;
    if (part) {
        int shift = 0;
        while ((w.range >> shift) > 256ull) ++shift;
        SDC_TRY(part_keys.alloc((size_t)n * 8, s));
        SDC_TRY(part_vals.alloc((size_t)n * 8, s));
        SDC_TRY(part_begin.alloc(257 * 8, s));
        SDC_TRY(range_partition_pass(keys, w.lo, shift, part_keys.p, 8, cols.data[0], part_vals.p, n, false,
                                     part_begin.as<unsigned long long>(), s));
        k = part_keys.p;
        v = part_vals.p;
        vec_ok = 1;
    }
    {
        const int64_t ntiles = ceil_div(n, kGbTile);
        int64_t grid = (int64_t)sms * 8;
        if (grid > ntiles) grid = ntiles;
        gb_window_kernel<<<(unsigned)grid, kGbThreads, 0, s>>>(reinterpret_cast<const unsigned long long*>(k),
                                                             reinterpret_cast<const unsigned long long*>(v), n, vec_ok, w, need,
                                                             mm.as<unsigned long long>() + 2);
        SDC_CHECK_LAUNCH();
    }
    const int64_t nt = ceil_div((int64_t)w.range, kWinTile);
    Scratch tcnt;
    SDC_TRY(tcnt.alloc((size_t)(nt + 1) * 8, s));
    gb_window_count_kernel<<<(unsigned)nt, 256, 0, s>>>(w, need, tcnt.as<unsigned long long>());
    SDC_CHECK_LAUNCH();
    scan_tiles_kernel<<<1, kScanTilesThreads, 0, s>>>(tcnt.as<unsigned long long>(), nt);
    SDC_CHECK_LAUNCH();
    unsigned long long total = 0;
    SDC_CUDA_TRY(cudaMemcpyAsync(&total, tcnt.as<unsigned long long>() + nt, 8, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    const int64_t ng = (int64_t)total;
    res->ngroups = ng;
    res->ngcap = ng;
    *done = true;
    if (ng == 0) return SDCB200_OK;
    SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->keys), (size_t)ng * 8, s));
    SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->vals), (size_t)ng * 8 * (nagg + 1), s));
    gb_window_write_kernel<<<(unsigned)nt, 256, 0, s>>>(w, need, agg_mask, tcnt.as<unsigned long long>(), ng, res->keys, res->vals);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

// pend != nullptr: the caller can finish later -- when the FIRST attempt is the one-kernel low-cardinality path and needs no
// second pass, the function returns right after the launch with *pend filled (sdcb200_groupby_end does the rest)
template <bool KEY_F64>
int32_t run_groupby(const void* keys, int64_t n, Cols cols, unsigned agg_mask, bool sort, long long ddof,
                    int key_dtype, int key_mode, int64_t expected_groups, cudaStream_t s, GroupbyResult* res,
                    PendingSmall** pend = nullptr) {
    const int need = need_for(agg_mask, sort);
    const bool need_ssd = (agg_mask & (AGG_VAR | AGG_STD)) != 0;
    const int sms = sm_count();
    const int nagg = __builtin_popcount(agg_mask);
    unsigned is_f64_mask = 0;
    for (int c = 0; c < cols.ncols; ++c) if (cols.is_f64[c]) is_f64_mask |= 1u << c;
    cols.vec_ok = (reinterpret_cast<uintptr_t>(keys) & 15) == 0;
    for (int c = 0; c < cols.ncols; ++c) if (reinterpret_cast<uintptr_t>(cols.data[c]) & 15) cols.vec_ok = 0;

    // plan: shared-memory pre-aggregation + small replicated table when the cardinality is (or may be) low;
    // a single global table sized from the hint while it stays around the size of L2; beyond that (hint or
    // overflow) the sort-and-segment path, whose cost does not depend on the number of groups
    const unsigned long long worst = pow2_at_least(2ull * (unsigned long long)n, 64);
    const bool big = n >= sort_path_rows();
    if (agg_mask & (AGG_PROD | AGG_MEDIAN))      // only the sort-and-segment path computes these
        return run_groupby_sorted<KEY_F64>(keys, n, cols, agg_mask, sort, ddof, key_dtype, key_mode, s, res);
    Attempt plan[3];
    int nattempts = 0;
    if (big && expected_groups > sort_path_groups()) {
        // Tens of millions of groups: a HASHED table costs two random HBM bursts per row and loses to sorting; int64 keys
        // that turn out to fill a dense range (what a ROUTE_MOD shuffle delivers, sdc_b200/distributed.py) take the
        // direct-addressed window path instead (run_groupby_window below), everything else the sort-and-segment path.
        if (!KEY_F64 && key_mode == KEYS_PLAIN && expected_groups <= dense_try_groups() && cols.ncols == 1 && cols.is_f64[0] &&
            !(need & NEED_FIRST) && !need_ssd && n < 0xffffffffll) {
            bool done = false;
            SDC_TRY(run_groupby_window(keys, n, cols, agg_mask, expected_groups, s, res, &done));
            if (done) return SDCB200_OK;
        }
        return run_groupby_sorted<KEY_F64>(keys, n, cols, agg_mask, sort, ddof, key_dtype, key_mode, s, res);
    } else if (expected_groups <= 0) {
        if (n >= 4096) plan[nattempts++] = {true, 8192, kSmallReplicas};
        if ((1ull << 22) < worst) plan[nattempts++] = {false, 1ull << 22, 1};
    } else if (expected_groups <= 4096 && n >= 4096) {
        plan[nattempts++] = {true, pow2_at_least(2ull * (unsigned long long)expected_groups, 1024), kSmallReplicas};
    } else if (key_mode == KEYS_GID) {
        plan[nattempts++] = {false, pow2_at_least((unsigned long long)expected_groups, 1024), 1};   // direct: one slot per id
    } else {
        const unsigned long long c = pow2_at_least(2ull * (unsigned long long)expected_groups, 1024);
        if (c < worst) plan[nattempts++] = {false, c, 1};
    }
    if (!big && key_mode != KEYS_GID)
        plan[nattempts++] = {false, worst, 1};      // small inputs: the worst-case table is cheap

    int allow_direct = 1, allow_lean_hash = 1;
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
        auto add = [&](void** dst, unsigned long long init, unsigned long long len) -> int32_t {
            SDC_TRY(mem.get(&p, len * 8, s));
            *dst = p;
            ip.ptr[ip.count] = reinterpret_cast<unsigned long long*>(p);
            ip.len[ip.count] = len;
            ip.val[ip.count++] = init;
            return SDCB200_OK;
        };
        SDC_TRY(add(reinterpret_cast<void**>(&t.keys), kEmptyKey(), per));
        if (need & NEED_FIRST) SDC_TRY(add(reinterpret_cast<void**>(&t.first), ~0ull, slots_n));
        for (int c = 0; c < cols.ncols; ++c) {
            if (need & NEED_SUM) SDC_TRY(add(&t.sum[c], 0ull, slots_n));
            if (need & NEED_CNT) SDC_TRY(add(reinterpret_cast<void**>(&t.cnt[c]), 0ull, slots_n));
            if (need & NEED_MIN) SDC_TRY(add(reinterpret_cast<void**>(&t.mn[c]), ~0ull, slots_n));
            if (need & NEED_MAX) SDC_TRY(add(reinterpret_cast<void**>(&t.mx[c]), 0ull, slots_n));
            if (need_ssd) SDC_TRY(add(reinterpret_cast<void**>(&t.ssd[c]), 0ull, per));
        }
        // one 256-byte control block: [0] occupied, [8] overflow flag, [16] groups written, [17] CTAs done,
        // [18] direct lo, [19] direct range, [20] compaction counter
        SDC_TRY(mem.get(&p, 256, s));
        ip.ptr[ip.count] = reinterpret_cast<unsigned long long*>(p);
        ip.len[ip.count] = 32;
        ip.val[ip.count++] = 0ull;
        t.occupied = reinterpret_cast<unsigned long long*>(p);
        t.overflow = reinterpret_cast<int*>(reinterpret_cast<char*>(p) + 64);
        unsigned long long* dmeta = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(p) + 128);
        {
            unsigned gx = (unsigned)(ceil_div((int64_t)slots_n, 1024) < (int64_t)sms * 4 ? ceil_div((int64_t)slots_n, 1024) : (int64_t)sms * 4);
            if (gx == 0) gx = 1;
            dim3 grid(gx, ip.count < 16 ? ip.count : 16);
            gb_init_kernel<<<grid, 256, 0, s>>>(ip);
            SDC_CHECK_LAUNCH();
        }
        unsigned long long hstack[32];
        unsigned long long* hmeta = pinned_meta();
        if (!hmeta) hmeta = hstack;
        // asynchronous call: its own pinned slot, so that several calls can be in flight
        unsigned long long* aslot = nullptr;
        if (pend && attempt == 0 && at.smem && !need_ssd && gb_env_int("SDCB200_GB_MAPPED", 1)) aslot = g_slots.take();
        unsigned long long dlo = 0, drange = 0;
        if (at.smem) {
            // ---- low cardinality: one kernel aggregates, folds and (unless var / std) writes the result
            // one float64 column with a common aggregate set: compile-time specialised loop
            int spec = -1;
            if (cols.ncols == 1 && cols.is_f64[0] && (need == 1 || need == 2 || need == 3 || need == 15)) spec = need;
            void (*kern)(const void*, int64_t, Cols, Table, int, SmemLayout, SmallFin) = gb_smem_kernel<KEY_F64, -1>;
            if (spec == 1) kern = gb_smem_kernel<KEY_F64, 1>;
            else if (spec == 2) kern = gb_smem_kernel<KEY_F64, 2>;
            else if (spec == 3) kern = gb_smem_kernel<KEY_F64, 3>;
            else if (spec == 15) kern = gb_smem_kernel<KEY_F64, 15>;
            // int64 keys / group ids: the instantiation with the lean loop for a direct window (it carries the generic
            // loop as well, for sampled keys that do not fit one)
            // Keys that have no such window (sparse int64 keys, every float64 key column) take the same instantiation's
            // hashed lean loop until it reports a key it could not place (overflow 3).
            bool lean = false;
            const bool want_direct = !KEY_F64 && allow_direct;
            const bool want_hash = allow_lean_hash && key_mode == KEYS_PLAIN && gb_env_int("SDCB200_GB_LEAN_HASH", 1) != 0;
            if (spec >= 0 && (want_direct || want_hash) && gb_lean_enabled()) {
                lean = true;
                if (spec == 1) kern = gb_smem_kernel<KEY_F64, 1, true>;
                else if (spec == 2) kern = gb_smem_kernel<KEY_F64, 2, true>;
                else if (spec == 3) kern = gb_smem_kernel<KEY_F64, 3, true>;
                else if (spec == 15) kern = gb_smem_kernel<KEY_F64, 15, true>;
            }
            const int rf_max = gb_env_int("SDCB200_GB_RF", 32);
            // the lean loop fills whatever the table holds beyond the key window with lane-indexed copies of it
            int want_slots = (int)(2 * (expected_groups > 0 ? expected_groups : 2048));
            if (lean && rf_max > 1) want_slots = 4096;
            SmemLayout lay = make_layout(cols.ncols, need, want_slots, 100 * 1024, t.cap);
            {   // function attributes are per device and stick: set once per (instantiation, device)
                static std::mutex mu;
                static std::set<std::pair<const void*, int>> done;
                int dev = 0;
                SDC_CUDA_TRY(cudaGetDevice(&dev));
                std::lock_guard<std::mutex> lock(mu);
                if (done.insert({reinterpret_cast<const void*>(kern), dev}).second)
                    SDC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            }
            const int64_t ntiles = ceil_div(n, (int64_t)kSmThreads * kRows);
            int64_t grid = (int64_t)sms * 2;
            if (grid > ntiles) grid = ntiles;
            SmallFin fin;
            memset(&fin, 0, sizeof(fin));
            const int64_t ngcap = (int64_t)per;
            if (!need_ssd) {
                // one allocation: the key row, then the value rows
                SDC_TRY(dev_alloc(reinterpret_cast<void**>(&res->keys), (size_t)ngcap * 8 * (nagg * cols.ncols + 2), s));
                res->vals = res->keys + ngcap;
                res->one_block = true;
                res->ngcap = ngcap;
            }
            fin.out_keys = res->keys;
            fin.out_vals = res->vals;
            fin.ngcap = ngcap;
            fin.agg_mask = agg_mask;
            fin.is_f64_mask = is_f64_mask;
            fin.ddof = ddof;
            fin.by_first = sort ? 0 : 1;
            fin.key_f64 = KEY_F64 ? 1 : 0;
            fin.write_out = need_ssd ? 0 : 1;
            fin.rf_max = rf_max;
            fin.hot = gb_env_int("SDCB200_GB_HOT", 1);
            // (a hint that already says the groups cannot fit the shared table keeps the generic loop and its global fallback)
            fin.lean_hash = lean && want_hash && (expected_groups <= 0 || 2 * expected_groups <= (int64_t)lay.scap) ? 1 : 0;
            fin.dh.mode = key_mode;
            fin.dh.allow = allow_direct;
            fin.dh.range = key_mode == KEYS_GID ? (unsigned long long)expected_groups : 0ull;
            fin.dh.out = dmeta + 2;
            fin.meta = dmeta;
            fin.host_meta = (hmeta != hstack && gb_env_int("SDCB200_GB_MAPPED", 1)) ? pinned_meta_device() : nullptr;
            if (aslot) {
                void* d = nullptr;
                if (cudaHostGetDevicePointer(&d, aslot, 0) == cudaSuccess) fin.host_meta = reinterpret_cast<unsigned long long*>(d);
                else { cudaGetLastError(); g_slots.give(aslot); aslot = nullptr; }
            }
            if (aslot) aslot[8] = ~0ull;            // "not published": the kernel overwrites the whole block
            kern<<<(unsigned)grid, kSmThreads, lay.alloc_bytes, s>>>(keys, n, cols, t, need, lay, fin);
            SDC_CHECK_LAUNCH();
            if (aslot) {
                PendingSmall* ps = new PendingSmall();
                ps->slot = aslot;
                ps->table = mem.detach();
                ps->stream = s;
                ps->lean = lean;
                ps->lean_hash = fin.lean_hash != 0;
                ps->keys = keys; ps->n = n; ps->cols = cols; ps->agg_mask = agg_mask; ps->sort = sort; ps->ddof = ddof;
                ps->key_dtype = key_dtype; ps->key_mode = key_mode; ps->expected_groups = expected_groups;
                *pend = ps;
                return SDCB200_OK;
            }
            if (!fin.host_meta) SDC_CUDA_TRY(cudaMemcpyAsync(hmeta, t.occupied, 256, cudaMemcpyDeviceToHost, s));
            SDC_CUDA_TRY(cudaStreamSynchronize(s));
            const int ovf = (int)(hmeta[8] & 0xffffffffull);
            if (ovf) {
                if (res->keys) dev_free(res->keys, s);
                if (res->vals && !res->one_block) dev_free(res->vals, s);
                res->keys = res->vals = nullptr;
                res->one_block = false;
                if (ovf == 2 && allow_direct) { allow_direct = 0; --attempt; continue; }   // same table, hashing this time
                if (ovf == 3 && allow_lean_hash) { allow_lean_hash = 0; --attempt; continue; }   // same table, generic loop
                if (attempt + 1 < nattempts || big) continue;
                set_error("groupby: hash table overflow at capacity %llu", t.cap);
                return SDCB200_ERR_CAPACITY;
            }
            if (!need_ssd) {
                res->ngroups = (int64_t)hmeta[16];
                res->lean_direct = lean && (hmeta[19] != 0ull || fin.lean_hash != 0);
                return SDCB200_OK;
            }
            dlo = hmeta[18];
            drange = hmeta[19];
            hmeta[0] = hmeta[16];                    // group count for the second-pass path below
        } else {
            const int64_t ntiles = ceil_div(n, kGbTile);
            int64_t grid = (int64_t)sms * 8;
            if (grid > ntiles) grid = ntiles;
            DirectHint dh;
            dh.mode = key_mode;
            dh.allow = allow_direct;
            dh.range = key_mode == KEYS_GID ? (unsigned long long)expected_groups : 0ull;
            dh.out = dmeta + 2;
            gb_global_kernel<KEY_F64><<<(unsigned)grid, kGbThreads, 0, s>>>(keys, n, cols, t, need, dh);
            SDC_CHECK_LAUNCH();
            SDC_CUDA_TRY(cudaMemcpyAsync(hmeta, t.occupied, 256, cudaMemcpyDeviceToHost, s));
            SDC_CUDA_TRY(cudaStreamSynchronize(s));
            const int ovf = (int)(hmeta[8] & 0xffffffffull);
            if (ovf) {
                if (ovf == 2 && allow_direct && key_mode != KEYS_GID) { allow_direct = 0; --attempt; continue; }
                if (ovf == 2 && key_mode == KEYS_GID) { set_error("groupby: a group id lies outside [0, %lld)", (long long)expected_groups); return SDCB200_ERR_ARG; }
                if (attempt + 1 < nattempts || big) continue;
                set_error("groupby: hash table overflow at capacity %llu", t.cap);
                return SDCB200_ERR_CAPACITY;
            }
            dlo = hmeta[18];
            drange = hmeta[19];
        }
        // large table (or var / std, which need a second pass over the rows once the means are known)
        if (need_ssd) {
            const int64_t ntiles = ceil_div(n, kGbTile);
            int64_t grid = (int64_t)sms * 8;
            if (grid > ntiles) grid = ntiles;
            gb_ssd_kernel<KEY_F64><<<(unsigned)grid, kGbThreads, 0, s>>>(keys, n, cols, t, dlo, drange);
            SDC_CHECK_LAUNCH();
        }
        // a direct-addressed table counts its groups while compacting; a hashed one counted them on insert
        const bool count_late = !at.smem && drange != 0;
        int64_t ng = count_late ? (int64_t)(n < (int64_t)per ? n : (int64_t)per) : (int64_t)hmeta[0];
        if (ng == 0) { res->ngroups = 0; res->ngcap = 0; return SDCB200_OK; }
        Scratch slots, sortkeys, order;
        SDC_TRY(slots.alloc((size_t)ng * 8, s));
        SDC_TRY(sortkeys.alloc((size_t)ng * 8, s));
        const int cgrid = (int)(ceil_div((int64_t)per, 256) < (int64_t)sms * 8 ? ceil_div((int64_t)per, 256) : (int64_t)sms * 8);
        gb_compact_kernel<<<cgrid, 256, 0, s>>>(t, slots.as<unsigned long long>(), sortkeys.as<unsigned long long>(),
                                                sort ? 0 : 1, dmeta + 4);
        SDC_CHECK_LAUNCH();
        if (count_late) {
            SDC_CUDA_TRY(cudaMemcpyAsync(hmeta, dmeta + 4, 8, cudaMemcpyDeviceToHost, s));
            SDC_CUDA_TRY(cudaStreamSynchronize(s));
            ng = (int64_t)hmeta[0];
        }
        res->ngroups = ng;
        res->ngcap = ng;
        if (ng == 0) return SDCB200_OK;
        SDC_TRY(order.alloc((size_t)ng * 8, s));
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
    if (big) return run_groupby_sorted<KEY_F64>(keys, n, cols, agg_mask, sort, ddof, key_dtype, key_mode, s, res);
    return SDCB200_ERR_CAPACITY;
}

void release_buffers(GroupbyResult* r, cudaStream_t s) {
    if (r->keys) dev_free(r->keys, s);
    if (r->vals && !r->one_block) dev_free(r->vals, s);
    r->keys = r->vals = nullptr;
    r->one_block = false;
}

// Several float64 columns over low-cardinality int64 keys: the one-column lean loop (lane-indexed window copies, hot
// keys in registers, ~20 instructions per row) once per column beats the generic loop over all columns at once even
// though it re-reads the keys -- 2 columns x 100 M rows: 0.82 against 0.89 ms on uniform keys, 1.1 against 4.3 ms on
// Zipf keys.  The first column tells whether the keys took that path; if not, the generic call runs as before.
int32_t run_groupby_int_keys(const void* keys, int64_t n, const Cols& cols, unsigned agg_mask, bool sort, long long ddof,
                             int key_dtype, int key_mode, int64_t expected_groups, cudaStream_t s, GroupbyResult* res) {
    const int need = need_for(agg_mask, sort);
    bool split = cols.ncols >= 2 && n >= (int64_t(1) << 22) && expected_groups <= 4096 &&
                 (need == 1 || need == 2 || need == 3 || need == 15) &&
                 !(agg_mask & (AGG_VAR | AGG_STD | AGG_PROD | AGG_MEDIAN)) && gb_lean_enabled() &&
                 gb_env_int("SDCB200_GB_SPLIT_COLS", 1) != 0;
    for (int c = 0; c < cols.ncols && split; ++c) split = cols.is_f64[c] != 0;
    if (split) {
        const int nagg = __builtin_popcount(agg_mask);
        int64_t ngcap = 0, ng = 0;
        for (int c = 0; c < cols.ncols; ++c) {
            Cols one;
            memset(&one, 0, sizeof(one));
            one.ncols = 1;
            one.data[0] = cols.data[c];
            one.is_f64[0] = 1;
            GroupbyResult sub;
            const int32_t rc = run_groupby<false>(keys, n, one, agg_mask, sort, ddof, key_dtype, key_mode, expected_groups, s, &sub);
            if (rc != SDCB200_OK) { release_buffers(&sub, s); release_buffers(res, s); return rc; }
            if (!sub.lean_direct || (c > 0 && (sub.ngroups != ng || sub.ngcap != ngcap))) {
                release_buffers(&sub, s);
                release_buffers(res, s);
                split = false;                                   // the generic call below
                break;
            }
            if (c == 0) {
                ng = sub.ngroups;
                ngcap = sub.ngcap;
                const int32_t arc = dev_alloc(reinterpret_cast<void**>(&res->keys), (size_t)ngcap * 8 * (nagg * cols.ncols + 2), s);
                if (arc != SDCB200_OK) { release_buffers(&sub, s); return arc; }
                res->vals = res->keys + ngcap;
                res->one_block = true;
            }
            // the sub-result holds one row per aggregate; here row (rank * ncols + c)
            cudaError_t ce = cudaSuccess;
            if (ng && c == 0) ce = cudaMemcpyAsync(res->keys, sub.keys, (size_t)ng * 8, cudaMemcpyDeviceToDevice, s);
            if (ng && ce == cudaSuccess)
                ce = cudaMemcpy2DAsync(res->vals + (size_t)c * ngcap, (size_t)cols.ncols * ngcap * 8, sub.vals,
                                       (size_t)ngcap * 8, (size_t)ng * 8, (size_t)nagg, cudaMemcpyDeviceToDevice, s);
            release_buffers(&sub, s);
            if (ce != cudaSuccess) {
                release_buffers(res, s);
                set_error("groupby: %s", cudaGetErrorString(ce));
                return SDCB200_ERR_CUDA;
            }
        }
        if (split) {
            res->ngroups = ng;
            res->ngcap = ngcap;
            res->lean_direct = true;
            return SDCB200_OK;
        }
    }
    return run_groupby<false>(keys, n, cols, agg_mask, sort, ddof, key_dtype, key_mode, expected_groups, s, res);
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int32_t sdcb200_groupby(int32_t key_dtype, const void* keys, int64_t n, int32_t ncols, const void* const* cols,
                        const int32_t* col_dtypes, uint32_t agg_mask, int32_t sort, int64_t ddof,
                        int64_t expected_groups, void* stream, int64_t* handle_out, int64_t* ngroups_out) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64 || key_dtype == SDCB200_GID, "key dtype must be I64, F64 or GID");
    int key_mode = KEYS_PLAIN;
    if (key_dtype == SDCB200_GID) {
        SDC_ARG_CHECK(expected_groups > 0, "GID keys: expected_groups is the id range and must be > 0");
        key_mode = KEYS_GID;
        key_dtype = SDCB200_I64;
    }
    SDC_ARG_CHECK(n >= 0 && ncols >= 0 && ncols <= kMaxCols, "n / ncols");
    SDC_ARG_CHECK(agg_mask != 0 && (agg_mask & ~0x1ffu) == 0, "agg mask");
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
    res->stream = s;
    res->key_dtype = key_dtype;
    res->ncols = ncols;
    res->agg_mask = agg_mask;
    int32_t rc = SDCB200_OK;
    if (n > 0) {
        rc = key_dtype == SDCB200_F64
                 ? run_groupby<true>(keys, n, c, agg_mask, sort != 0, ddof, key_dtype, KEYS_PLAIN, expected_groups, s, res)
                 : run_groupby_int_keys(keys, n, c, agg_mask, sort != 0, ddof, key_dtype, key_mode, expected_groups, s, res);
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

// Asynchronous form of sdcb200_groupby for pipelines of small aggregates (C1: 10 M rows take ~50 us of kernel time, a
// blocking call adds a stream synchronisation and the host's launch path to every one of them).  _begin enqueues the work
// and returns; _end waits for the stream and yields the group count.  When the request takes the one-kernel
// low-cardinality path (int64 / float64 keys, no var / std / prod / median, at most one value column, a group hint
// <= 4096 or none) nothing blocks in _begin: the kernel's last CTA publishes the control block to mapped pinned memory.
// Anything else runs to completion inside _begin.  THE INPUT COLUMNS MUST STAY VALID UNTIL _end: an optimistic attempt
// that overflowed (a key outside the sampled window, more groups than hinted) is rerun there.
int32_t sdcb200_groupby_begin(int32_t key_dtype, const void* keys, int64_t n, int32_t ncols, const void* const* cols,
                              const int32_t* col_dtypes, uint32_t agg_mask, int32_t sort, int64_t ddof, int64_t expected_groups,
                              void* stream, int64_t* handle_out) {
    const bool simple = ncols <= 1 && !(agg_mask & (AGG_VAR | AGG_STD | AGG_PROD | AGG_MEDIAN)) && n >= 4096 &&
                        expected_groups <= 4096 && (key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64);
    if (!simple) {
        int64_t ng = 0;
        return sdcb200_groupby(key_dtype, keys, n, ncols, cols, col_dtypes, agg_mask, sort, ddof, expected_groups, stream, handle_out, &ng);
    }
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(agg_mask != 0 && (agg_mask & ~0x1ffu) == 0, "agg mask");
    SDC_ARG_CHECK(handle_out != nullptr && keys != nullptr, "null pointer");
    Cols c;
    memset(&c, 0, sizeof(c));
    c.ncols = ncols;
    for (int i = 0; i < ncols; ++i) {
        SDC_ARG_CHECK(col_dtypes[i] == SDCB200_I64 || col_dtypes[i] == SDCB200_F64, "column dtype must be I64 or F64");
        SDC_ARG_CHECK(cols[i] != nullptr, "null column");
        c.data[i] = cols[i];
        c.is_f64[i] = col_dtypes[i] == SDCB200_F64;
    }
    GroupbyResult* res = new GroupbyResult();
    res->stream = s;
    res->key_dtype = key_dtype;
    res->ncols = ncols;
    res->agg_mask = agg_mask;
    PendingSmall* pend = nullptr;
    const int32_t rc = key_dtype == SDCB200_F64
                           ? run_groupby<true>(keys, n, c, agg_mask, sort != 0, ddof, key_dtype, KEYS_PLAIN, expected_groups, s, res, &pend)
                           : run_groupby<false>(keys, n, c, agg_mask, sort != 0, ddof, key_dtype, KEYS_PLAIN, expected_groups, s, res, &pend);
    if (rc != SDCB200_OK) {
        free_result(res);
        return rc;
    }
    res->pending = pend;
    std::lock_guard<std::mutex> lk(g_mu);
    const int64_t h = g_next_handle++;
    g_results[h] = res;
    *handle_out = h;
    return SDCB200_OK;
}

int32_t sdcb200_groupby_end(int64_t handle, int64_t* ngroups_out) {
    GroupbyResult* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_results.find(handle);
        if (it != g_results.end()) r = it->second;
    }
    if (!r) { set_error("groupby: unknown handle %lld", (long long)handle); return SDCB200_ERR_HANDLE; }
    SDC_ARG_CHECK(ngroups_out != nullptr, "null out pointer");
    if (PendingSmall* ps = r->pending) {
        r->pending = nullptr;
        const cudaError_t e = cudaStreamSynchronize(ps->stream);
        const int ovf = (int)(ps->slot[8] & 0xffffffffull);
        const int64_t ng = (int64_t)ps->slot[16];
        const bool direct = ps->slot[19] != 0ull;
        g_slots.give(ps->slot);
        dev_free(ps->table, ps->stream);
        int32_t rc = SDCB200_OK;
        if (e != cudaSuccess) {
            set_error("groupby: %s", cudaGetErrorString(e));
            rc = SDCB200_ERR_CUDA;
        } else if (ovf) {
            // the optimistic attempt did not fit: the blocking path, from the start (it falls through to hashing / larger tables)
            release_buffers(r, ps->stream);
            rc = ps->key_dtype == SDCB200_F64
                     ? run_groupby<true>(ps->keys, ps->n, ps->cols, ps->agg_mask, ps->sort, ps->ddof, ps->key_dtype, ps->key_mode,
                                         ps->expected_groups, ps->stream, r)
                     : run_groupby<false>(ps->keys, ps->n, ps->cols, ps->agg_mask, ps->sort, ps->ddof, ps->key_dtype, ps->key_mode,
                                          ps->expected_groups, ps->stream, r);
        } else {
            r->ngroups = ng;
            r->lean_direct = ps->lean && (direct || ps->lean_hash);
        }
        delete ps;
        if (rc != SDCB200_OK) return rc;
    }
    *ngroups_out = r->ngroups;
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

int32_t sdcb200_groupby_result(int64_t handle, void** keys, void** vals, int64_t* stride) {
    GroupbyResult* r = find_result(handle);
    if (!r) { set_error("groupby: unknown handle %lld", (long long)handle); return SDCB200_ERR_HANDLE; }
    SDC_ARG_CHECK(keys != nullptr && vals != nullptr && stride != nullptr, "null out pointer");
    *keys = r->keys;
    *vals = r->vals;
    *stride = r->ngcap;
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

extern "C" int32_t sdcb200_combine_ids(const int64_t* a, const int64_t* b, int64_t nb, int64_t invalid, int64_t* out, int64_t n,
                                       void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(n >= 0 && nb >= 0, "n / nb");
    if (n == 0) return SDCB200_OK;
    SDC_ARG_CHECK(a != nullptr && b != nullptr && out != nullptr, "null column");
    int64_t grid = ceil_div(n, 256 * 4);
    if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
    combine_ids_kernel<<<(unsigned)grid, 256, 0, s>>>(reinterpret_cast<const long long*>(a), reinterpret_cast<const long long*>(b),
                                                      (long long)nb, (long long)invalid, reinterpret_cast<long long*>(out), n);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

// Host-buffer form (what a jitted SDC function holds: NumPy / NRT memory): the key and value columns go to the
// device on two copy streams, the aggregate runs, the inputs are released; the result stays in the handle and is
// read with sdcb200_groupby_keys / _values (dst_is_host = 1).  The copies dominate (16 B per row over PCIe).
extern "C" int32_t sdcb200_host_groupby(int32_t key_dtype, const void* keys, int64_t n, int32_t ncols, const void* const* cols,
                                        const int32_t* col_dtypes, uint32_t agg_mask, int32_t sort, int64_t ddof,
                                        int64_t expected_groups, int64_t* handle_out, int64_t* ngroups_out) {
    SDC_ARG_CHECK(key_dtype == SDCB200_I64 || key_dtype == SDCB200_F64, "key dtype must be I64 or F64");
    SDC_ARG_CHECK(n >= 0 && ncols >= 0 && ncols <= kMaxCols, "n / ncols");
    SDC_ARG_CHECK(n == 0 || keys != nullptr, "null keys");
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ev = nullptr;
    void* dkeys = nullptr;
    void* dcols[kMaxCols];
    for (int i = 0; i < kMaxCols; ++i) dcols[i] = nullptr;
    auto body = [&]() -> int32_t {
        for (int i = 0; i < 2; ++i) SDC_CUDA_TRY(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        SDC_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        const size_t bytes = (size_t)(n > 0 ? n : 1) * 8;
        SDC_TRY(dev_alloc(&dkeys, bytes, st[0]));
        if (n > 0) SDC_CUDA_TRY(cudaMemcpyAsync(dkeys, keys, (size_t)n * 8, cudaMemcpyHostToDevice, st[0]));
        for (int i = 0; i < ncols; ++i) {
            SDC_ARG_CHECK(n == 0 || cols[i] != nullptr, "null column");
            cudaStream_t cs = st[(i + 1) & 1];
            SDC_TRY(dev_alloc(&dcols[i], bytes, cs));
            if (n > 0) SDC_CUDA_TRY(cudaMemcpyAsync(dcols[i], cols[i], (size_t)n * 8, cudaMemcpyHostToDevice, cs));
        }
        SDC_CUDA_TRY(cudaEventRecord(ev, st[1]));
        SDC_CUDA_TRY(cudaStreamWaitEvent(st[0], ev, 0));
        SDC_TRY(sdcb200_groupby(key_dtype, dkeys, n, ncols, dcols, col_dtypes, agg_mask, sort, ddof, expected_groups, st[0],
                                handle_out, ngroups_out));
        return SDCB200_OK;
    };
    const int32_t rc = body();
    if (rc == SDCB200_OK) {
        // the result was produced on st[0], which is synchronised and destroyed below: later frees go to the default stream
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_results.find(*handle_out);
        if (it != g_results.end()) it->second->stream = nullptr;
    }
    if (dkeys) dev_free(dkeys, st[0]);
    for (int i = 0; i < ncols; ++i)
        if (dcols[i]) dev_free(dcols[i], st[0]);
    for (int i = 0; i < 2; ++i)
        if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
    if (ev) cudaEventDestroy(ev);
    return rc;
}
