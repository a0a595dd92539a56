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

constexpr unsigned long long kEmpty = 0x7ff8dead0000beefull;   // a NaN payload: never a float key;
                                                               // an int64 key equal to it uses the spare slot
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

struct Table {
    unsigned long long* keys;       // [cap + 1]; slot cap = the spare slot for key == kEmpty
    unsigned long long* first;      // [cap + 1] first row of the group (NEED_FIRST)
    // per column c, arrays of cap + 1 entries (nullptr when not needed)
    void* sum[kMaxCols];            // double or long long
    unsigned long long* cnt[kMaxCols];
    unsigned long long* mn[kMaxCols];   // ordered u64
    unsigned long long* mx[kMaxCols];
    double* ssd[kMaxCols];          // second pass: sum of squared deviations
    unsigned long long cap;         // power of two
    unsigned long long limit;       // max occupied slots before "overflow"
    unsigned long long* occupied;   // device counter
    int* overflow;                  // device flag
};

struct Cols {
    const void* data[kMaxCols];
    int is_f64[kMaxCols];
    int ncols;
};

// key column access: canonical 64-bit key + validity
template <bool KEY_F64>
__device__ __forceinline__ bool load_key(const void* keys, int64_t row, unsigned long long& k) {
    unsigned long long b = reinterpret_cast<const unsigned long long*>(keys)[row];
    if (KEY_F64) {
        const unsigned long long mag = b & 0x7fffffffffffffffull;
        if (mag > 0x7ff0000000000000ull) return false;   // NaN key: row skipped
        if (mag == 0) b = 0;                             // -0.0 == +0.0
    }
    k = b;
    return true;
}

// find-or-insert in the global table; returns slot, or ~0 when the table overflowed
__device__ __forceinline__ unsigned long long global_upsert(const Table& t, unsigned long long key) {
    if (key == kEmpty) {
        if (atomicCAS(&t.keys[t.cap], kEmpty, key + 1) == kEmpty) atomicAdd(t.occupied, 1ull);
        return t.cap;
    }
    unsigned long long slot = mix64(key) & (t.cap - 1);
    for (unsigned long long probes = 0; probes <= t.cap; ++probes) {
        unsigned long long cur = t.keys[slot];
        if (cur == key) return slot;
        if (cur == kEmpty) {
            if (*(volatile int*)t.overflow) return ~0ull;
            const unsigned long long old = atomicCAS(&t.keys[slot], kEmpty, key);
            if (old == kEmpty) {
                if (atomicAdd(t.occupied, 1ull) + 1 > t.limit) { *(volatile int*)t.overflow = 1; return ~0ull; }
                return slot;
            }
            if (old == key) return slot;
        }
        slot = (slot + 1) & (t.cap - 1);
    }
    *(volatile int*)t.overflow = 1;
    return ~0ull;
}

// read-only lookup (second pass); the key is known to be present
__device__ __forceinline__ unsigned long long global_find(const Table& t, unsigned long long key) {
    if (key == kEmpty) return t.cap;
    unsigned long long slot = mix64(key) & (t.cap - 1);
    while (t.keys[slot] != key) slot = (slot + 1) & (t.cap - 1);
    return slot;
}

__device__ __forceinline__ void global_accumulate(const Table& t, const Cols& cols, int need, unsigned long long slot,
                                                  int64_t row) {
    if (need & NEED_FIRST) atomicMin(&t.first[slot], (unsigned long long)row);
#pragma unroll 1
    for (int c = 0; c < cols.ncols; ++c) {
        if (cols.is_f64[c]) {
            const double v = reinterpret_cast<const double*>(cols.data[c])[row];
            if (v != v) continue;                                  // skipna
            if (need & NEED_SUM) atomicAdd(reinterpret_cast<double*>(t.sum[c]) + slot, v);
            if (need & NEED_CNT) atomicAdd(&t.cnt[c][slot], 1ull);
            if (need & NEED_MIN) atomicMin(&t.mn[c][slot], f64_to_ordered(v));
            if (need & NEED_MAX) atomicMax(&t.mx[c][slot], f64_to_ordered(v));
        } else {
            const long long v = reinterpret_cast<const long long*>(cols.data[c])[row];
            if (need & NEED_SUM) atomicAdd(reinterpret_cast<unsigned long long*>(t.sum[c]) + slot, (unsigned long long)v);
            if (need & NEED_CNT) atomicAdd(&t.cnt[c][slot], 1ull);
            if (need & NEED_MIN) atomicMin(&t.mn[c][slot], i64_to_ordered(v));
            if (need & NEED_MAX) atomicMax(&t.mx[c][slot], i64_to_ordered(v));
        }
    }
}

// ------------------------------------------------------------------ high-cardinality path
template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_global_kernel(const void* __restrict__ keys, int64_t n, Cols cols,
                                                               Table t, int need) {
    const int64_t stride = (int64_t)gridDim.x * kGbThreads;
    for (int64_t row = (int64_t)blockIdx.x * kGbThreads + threadIdx.x; row < n; row += stride) {
        unsigned long long k;
        if (!load_key<KEY_F64>(keys, row, k)) continue;
        const unsigned long long slot = global_upsert(t, k);
        if (slot == ~0ull) return;
        global_accumulate(t, cols, need, slot, row);
    }
}

// ------------------------------------------------------------------ low-cardinality path
// Per-CTA shared-memory table of SCAP slots for up to kSmemCols columns; layout (SoA):
//   keys[SCAP] u64 | first[SCAP] u64 | per column: sum[SCAP] 8B, cnt[SCAP] u32, mn[SCAP] u64, mx[SCAP] u64
// (only the arrays `need` asks for are carved out).  Rows that cannot get a slot within
// kSmemProbes probes go to the global table directly.
constexpr int kSmemProbes = 16;

struct SmemLayout {
    int scap;
    int off_first;          // byte offsets from the table base; -1 = absent
    int off_sum[kMaxCols], off_cnt[kMaxCols], off_mn[kMaxCols], off_mx[kMaxCols];
    int bytes;
};

template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_smem_kernel(const void* __restrict__ keys, int64_t n, Cols cols, Table t,
                                                             int need, SmemLayout lay) {
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sm);
    const int scap = lay.scap;
    // init
    for (int i = threadIdx.x; i < scap; i += kGbThreads) {
        skeys[i] = kEmpty;
        if (need & NEED_FIRST) reinterpret_cast<unsigned long long*>(sm + lay.off_first)[i] = ~0ull;
        for (int c = 0; c < cols.ncols; ++c) {
            if (need & NEED_SUM) reinterpret_cast<unsigned long long*>(sm + lay.off_sum[c])[i] = 0ull;
            if (need & NEED_CNT) reinterpret_cast<unsigned*>(sm + lay.off_cnt[c])[i] = 0u;
            if (need & NEED_MIN) reinterpret_cast<unsigned long long*>(sm + lay.off_mn[c])[i] = ~0ull;
            if (need & NEED_MAX) reinterpret_cast<unsigned long long*>(sm + lay.off_mx[c])[i] = 0ull;
        }
    }
    __syncthreads();
    // contiguous row range per CTA keeps the loads coalesced and the first-row minima cheap
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = per * blockIdx.x;
    const int64_t hi = lo + per < n ? lo + per : n;
    for (int64_t row = lo + threadIdx.x; row < hi; row += kGbThreads) {
        unsigned long long k;
        if (!load_key<KEY_F64>(keys, row, k)) continue;
        int slot = -1;
        if (k != kEmpty) {
            unsigned s = (unsigned)mix64(k) & (unsigned)(scap - 1);
#pragma unroll 1
            for (int p = 0; p < kSmemProbes; ++p) {
                unsigned long long cur = skeys[s];
                if (cur == k) { slot = (int)s; break; }
                if (cur == kEmpty) {
                    const unsigned long long old = atomicCAS(&skeys[s], kEmpty, k);
                    if (old == kEmpty || old == k) { slot = (int)s; break; }
                }
                s = (s + 1) & (unsigned)(scap - 1);
            }
        }
        if (slot < 0) {     // shared table full around this key: aggregate in the global table
            const unsigned long long g = global_upsert(t, k);
            if (g != ~0ull) global_accumulate(t, cols, need, g, row);
            continue;
        }
        if (need & NEED_FIRST) atomicMin(reinterpret_cast<unsigned long long*>(sm + lay.off_first) + slot, (unsigned long long)row);
#pragma unroll 1
        for (int c = 0; c < cols.ncols; ++c) {
            if (cols.is_f64[c]) {
                const double v = reinterpret_cast<const double*>(cols.data[c])[row];
                if (v != v) continue;
                if (need & NEED_SUM) atomicAdd(reinterpret_cast<double*>(sm + lay.off_sum[c]) + slot, v);
                if (need & NEED_CNT) atomicAdd(reinterpret_cast<unsigned*>(sm + lay.off_cnt[c]) + slot, 1u);
                if (need & NEED_MIN) atomicMin(reinterpret_cast<unsigned long long*>(sm + lay.off_mn[c]) + slot, f64_to_ordered(v));
                if (need & NEED_MAX) atomicMax(reinterpret_cast<unsigned long long*>(sm + lay.off_mx[c]) + slot, f64_to_ordered(v));
            } else {
                const long long v = reinterpret_cast<const long long*>(cols.data[c])[row];
                if (need & NEED_SUM) atomicAdd(reinterpret_cast<unsigned long long*>(sm + lay.off_sum[c]) + slot, (unsigned long long)v);
                if (need & NEED_CNT) atomicAdd(reinterpret_cast<unsigned*>(sm + lay.off_cnt[c]) + slot, 1u);
                if (need & NEED_MIN) atomicMin(reinterpret_cast<unsigned long long*>(sm + lay.off_mn[c]) + slot, i64_to_ordered(v));
                if (need & NEED_MAX) atomicMax(reinterpret_cast<unsigned long long*>(sm + lay.off_mx[c]) + slot, i64_to_ordered(v));
            }
        }
    }
    __syncthreads();
    // flush the occupied shared slots into the global table
    for (int i = threadIdx.x; i < scap; i += kGbThreads) {
        const unsigned long long k = skeys[i];
        if (k == kEmpty) continue;
        const unsigned long long g = global_upsert(t, k);
        if (g == ~0ull) continue;
        if (need & NEED_FIRST) atomicMin(&t.first[g], reinterpret_cast<unsigned long long*>(sm + lay.off_first)[i]);
        for (int c = 0; c < cols.ncols; ++c) {
            unsigned cn = 1;
            if (need & NEED_CNT) {
                cn = reinterpret_cast<unsigned*>(sm + lay.off_cnt[c])[i];
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

// second pass for var / std: ssd[slot] += (x - mean)^2
template <bool KEY_F64>
__global__ void __launch_bounds__(kGbThreads) gb_ssd_kernel(const void* __restrict__ keys, int64_t n, Cols cols, Table t) {
    const int64_t stride = (int64_t)gridDim.x * kGbThreads;
    for (int64_t row = (int64_t)blockIdx.x * kGbThreads + threadIdx.x; row < n; row += stride) {
        unsigned long long k;
        if (!load_key<KEY_F64>(keys, row, k)) continue;
        const unsigned long long slot = global_find(t, k);
        for (int c = 0; c < cols.ncols; ++c) {
            double v, mean;
            const double cnt = (double)t.cnt[c][slot];
            if (cols.is_f64[c]) {
                v = reinterpret_cast<const double*>(cols.data[c])[row];
                if (v != v) continue;
                mean = reinterpret_cast<const double*>(t.sum[c])[slot] / cnt;
            } else {
                v = (double)reinterpret_cast<const long long*>(cols.data[c])[row];
                mean = (double)reinterpret_cast<const long long*>(t.sum[c])[slot] / cnt;
            }
            const double d = v - mean;
            atomicAdd(&t.ssd[c][slot], d * d);
        }
    }
}

__global__ void gb_fill_kernel(unsigned long long* p, unsigned long long v, unsigned long long n) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// compact occupied slots: slots[j] = slot id, gkeys[j] = key (or first row when by_first)
__global__ void gb_compact_kernel(Table t, unsigned long long* slots, unsigned long long* sortkeys, int by_first,
                                  unsigned long long* counter) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i <= t.cap;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = t.keys[i];
        if (k == kEmpty) continue;
        const unsigned long long j = atomicAdd(counter, 1ull);
        slots[j] = i;
        sortkeys[j] = by_first ? t.first[i] : (i == t.cap ? kEmpty : k);
    }
}

struct OutPtrs {
    void* keys;
    void* val[kMaxCols];   // one aggregate per launch
};

// gather one aggregate (agg) of every column in result order
__global__ void gb_finalize_kernel(Table t, Cols cols, const unsigned long long* __restrict__ slots,
                                   const int64_t* __restrict__ order, int64_t ng, unsigned agg, long long ddof,
                                   OutPtrs out, int write_keys) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ng; j += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long slot = slots[order[j]];
        if (write_keys) {
            const unsigned long long k = slot == t.cap ? kEmpty : t.keys[slot];
            reinterpret_cast<unsigned long long*>(out.keys)[j] = k;
        }
        for (int c = 0; c < cols.ncols; ++c) {
            if (!out.val[c]) continue;
            const unsigned long long cn = t.cnt[c] ? t.cnt[c][slot] : 0ull;
            const bool f = cols.is_f64[c] != 0;
            double* od = reinterpret_cast<double*>(out.val[c]);
            long long* oi = reinterpret_cast<long long*>(out.val[c]);
            switch (agg) {
                case AGG_SUM:
                    if (f) od[j] = reinterpret_cast<const double*>(t.sum[c])[slot];
                    else oi[j] = reinterpret_cast<const long long*>(t.sum[c])[slot];
                    break;
                case AGG_MEAN: {
                    const double s = f ? reinterpret_cast<const double*>(t.sum[c])[slot]
                                       : (double)reinterpret_cast<const long long*>(t.sum[c])[slot];
                    od[j] = cn ? s / (double)cn : nan;
                    break;
                }
                case AGG_MIN:
                case AGG_MAX: {
                    const unsigned long long o = agg == AGG_MIN ? t.mn[c][slot] : t.mx[c][slot];
                    if (f) od[j] = cn ? ordered_to_f64(o) : nan;
                    else oi[j] = (long long)(o ^ 0x8000000000000000ull);
                    break;
                }
                case AGG_COUNT:
                    oi[j] = (long long)cn;
                    break;
                case AGG_VAR:
                case AGG_STD: {
                    double v = nan;
                    if ((long long)cn > ddof) {
                        const double dn = (double)cn;
                        v = (t.ssd[c][slot] / dn) * dn / (double)((long long)cn - ddof);
                        if (agg == AGG_STD) v = sqrt(v);
                    }
                    od[j] = v;
                    break;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ host side
struct GroupbyResult {
    int key_dtype = 0;
    int ncols = 0;
    int64_t ngroups = 0;
    int col_is_f64[kMaxCols] = {0};
    unsigned agg_mask = 0;
    void* keys = nullptr;                          // [ngroups] 8-byte keys
    void* vals[7][kMaxCols] = {{nullptr}};         // [agg bit index][col] -> [ngroups]
    int device = 0;
};

std::mutex g_mu;
std::unordered_map<int64_t, GroupbyResult*> g_results;
int64_t g_next_handle = 1;

int agg_index(unsigned agg) {
    switch (agg) {
        case AGG_SUM: return 0; case AGG_MEAN: return 1; case AGG_MIN: return 2; case AGG_MAX: return 3;
        case AGG_COUNT: return 4; case AGG_VAR: return 5; case AGG_STD: return 6;
    }
    return -1;
}

void free_result(GroupbyResult* r) {
    if (!r) return;
    if (r->keys) cudaFree(r->keys);
    for (int a = 0; a < 7; ++a)
        for (int c = 0; c < kMaxCols; ++c)
            if (r->vals[a][c]) cudaFree(r->vals[a][c]);
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

SmemLayout make_layout(int ncols, int need, int budget_bytes) {
    SmemLayout L;
    int per_slot = 8 + ((need & NEED_FIRST) ? 8 : 0);
    per_slot += ncols * (((need & NEED_SUM) ? 8 : 0) + ((need & NEED_CNT) ? 4 : 0) + ((need & NEED_MIN) ? 8 : 0) +
                         ((need & NEED_MAX) ? 8 : 0));
    int scap = 4096;
    while (scap > 64 && scap * per_slot > budget_bytes) scap >>= 1;
    L.scap = scap;
    int off = scap * 8;
    L.off_first = -1;
    if (need & NEED_FIRST) { L.off_first = off; off += scap * 8; }
    for (int c = 0; c < kMaxCols; ++c) L.off_sum[c] = L.off_cnt[c] = L.off_mn[c] = L.off_mx[c] = -1;
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

struct TableMem {
    std::vector<Scratch*> blocks;
    ~TableMem() { for (auto* b : blocks) delete b; }
    int32_t get(void** p, size_t bytes, cudaStream_t s) {
        Scratch* b = new Scratch();
        blocks.push_back(b);
        SDC_TRY(b->alloc(bytes, s));
        *p = b->p;
        return SDCB200_OK;
    }
};

template <bool KEY_F64>
int32_t run_groupby(const void* keys, int64_t n, const Cols& cols, unsigned agg_mask, bool sort, long long ddof,
                    int key_dtype, int64_t expected_groups, cudaStream_t s, GroupbyResult* res) {
    const int need = need_for(agg_mask, sort);
    const bool need_ssd = (agg_mask & (AGG_VAR | AGG_STD)) != 0;
    const int sms = sm_count();
    // capacity attempts: a table that lives in L2 first, the worst case (every row its own group) second
    unsigned long long caps[2];
    int nattempts = 0;
    unsigned long long worst = 64;
    while (worst < 2ull * (unsigned long long)n) worst <<= 1;
    unsigned long long first_cap = 1ull << 20;
    if (expected_groups > 0) { first_cap = 64; while (first_cap < 2ull * (unsigned long long)expected_groups) first_cap <<= 1; }
    if (first_cap < worst) caps[nattempts++] = first_cap;
    caps[nattempts++] = worst;

    for (int attempt = 0; attempt < nattempts; ++attempt) {
        TableMem mem;
        Table t;
        memset(&t, 0, sizeof(t));
        t.cap = caps[attempt];
        t.limit = t.cap / 2 + 1;
        const unsigned long long slots_n = t.cap + 1;
        const int fill_grid = (int)(ceil_div((int64_t)slots_n, 256 * 8) < sms * 8 ? ceil_div((int64_t)slots_n, 256 * 8) : sms * 8);
        auto fill = [&](void* p, unsigned long long v) -> int32_t {
            gb_fill_kernel<<<fill_grid, 256, 0, s>>>(reinterpret_cast<unsigned long long*>(p), v, slots_n);
            SDC_CHECK_LAUNCH();
            return SDCB200_OK;
        };
        void* p;
        SDC_TRY(mem.get(&p, slots_n * 8, s)); t.keys = (unsigned long long*)p; SDC_TRY(fill(p, kEmpty));
        if (need & NEED_FIRST) { SDC_TRY(mem.get(&p, slots_n * 8, s)); t.first = (unsigned long long*)p; SDC_TRY(fill(p, ~0ull)); }
        for (int c = 0; c < cols.ncols; ++c) {
            if (need & NEED_SUM) { SDC_TRY(mem.get(&p, slots_n * 8, s)); t.sum[c] = p; SDC_CUDA_TRY(cudaMemsetAsync(p, 0, slots_n * 8, s)); }
            if (need & NEED_CNT) { SDC_TRY(mem.get(&p, slots_n * 8, s)); t.cnt[c] = (unsigned long long*)p; SDC_CUDA_TRY(cudaMemsetAsync(p, 0, slots_n * 8, s)); }
            if (need & NEED_MIN) { SDC_TRY(mem.get(&p, slots_n * 8, s)); t.mn[c] = (unsigned long long*)p; SDC_TRY(fill(p, ~0ull)); }
            if (need & NEED_MAX) { SDC_TRY(mem.get(&p, slots_n * 8, s)); t.mx[c] = (unsigned long long*)p; SDC_CUDA_TRY(cudaMemsetAsync(p, 0, slots_n * 8, s)); }
            if (need_ssd) { SDC_TRY(mem.get(&p, slots_n * 8, s)); t.ssd[c] = (double*)p; SDC_CUDA_TRY(cudaMemsetAsync(p, 0, slots_n * 8, s)); }
        }
        SDC_TRY(mem.get(&p, 64, s));
        SDC_CUDA_TRY(cudaMemsetAsync(p, 0, 64, s));
        t.occupied = (unsigned long long*)p;
        t.overflow = (int*)((char*)p + 8);
        unsigned long long* compact_counter = (unsigned long long*)((char*)p + 16);

        // shared-memory pre-aggregation unless the caller says the cardinality is high
        SmemLayout lay = make_layout(cols.ncols, need, 96 * 1024);
        const bool use_smem = (expected_groups <= 0 || expected_groups <= lay.scap) && n >= 4096;
        if (use_smem) {
            auto kern = gb_smem_kernel<KEY_F64>;
            static thread_local int configured = 0;
            if (lay.bytes > configured) {
                SDC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                configured = 200 * 1024;
            }
            int grid = sms * 2;
            if ((int64_t)grid * kGbThreads > n) grid = (int)ceil_div(n, kGbThreads);
            kern<<<grid, kGbThreads, lay.bytes, s>>>(keys, n, cols, t, need, lay);
        } else {
            int grid = (int)(ceil_div(n, kGbThreads) < (int64_t)sms * 8 ? ceil_div(n, kGbThreads) : (int64_t)sms * 8);
            gb_global_kernel<KEY_F64><<<grid, kGbThreads, 0, s>>>(keys, n, cols, t, need);
        }
        SDC_CHECK_LAUNCH();
        if (need_ssd) {
            int grid = (int)(ceil_div(n, kGbThreads) < (int64_t)sms * 8 ? ceil_div(n, kGbThreads) : (int64_t)sms * 8);
            gb_ssd_kernel<KEY_F64><<<grid, kGbThreads, 0, s>>>(keys, n, cols, t);
            SDC_CHECK_LAUNCH();
        }
        // compaction (the overflow flag is read back together with the group count)
        Scratch slots, sortkeys;
        SDC_TRY(slots.alloc((size_t)(t.limit + 2) * 8, s));
        SDC_TRY(sortkeys.alloc((size_t)(t.limit + 2) * 8, s));
        unsigned long long hmeta[3];
        SDC_CUDA_TRY(cudaMemcpyAsync(hmeta, t.occupied, 16, cudaMemcpyDeviceToHost, s));
        SDC_CUDA_TRY(cudaStreamSynchronize(s));
        const int overflow = (int)(hmeta[1] & 0xffffffffull);
        if (overflow) {
            if (attempt + 1 < nattempts) continue;
            set_error("groupby: hash table overflow at capacity %llu", t.cap);
            return SDCB200_ERR_CAPACITY;
        }
        const int64_t ng = (int64_t)hmeta[0];
        res->ngroups = ng;
        if (ng == 0) return SDCB200_OK;
        gb_compact_kernel<<<fill_grid, 256, 0, s>>>(t, slots.as<unsigned long long>(), sortkeys.as<unsigned long long>(),
                                                    sort ? 0 : 1, compact_counter);
        SDC_CHECK_LAUNCH();
        // order of the result rows
        Scratch order;
        SDC_TRY(order.alloc((size_t)ng * 8, s));
        SDC_TRY(argsort_device(sort ? key_dtype : SDCB200_U64, sortkeys.p, ng, true, order.as<int64_t>(), s));
        // outputs
        SDC_CUDA_TRY(cudaMalloc(&res->keys, (size_t)ng * 8));
        const int fgrid = (int)(ceil_div(ng, 256) < (int64_t)sms * 4 ? ceil_div(ng, 256) : (int64_t)sms * 4);
        bool keys_written = false;
        for (unsigned bit = 1; bit <= AGG_STD; bit <<= 1) {
            if (!(agg_mask & bit)) continue;
            OutPtrs out;
            memset(&out, 0, sizeof(out));
            out.keys = res->keys;
            for (int c = 0; c < cols.ncols; ++c) {
                SDC_CUDA_TRY(cudaMalloc(&res->vals[agg_index(bit)][c], (size_t)ng * 8));
                out.val[c] = res->vals[agg_index(bit)][c];
            }
            gb_finalize_kernel<<<fgrid, 256, 0, s>>>(t, cols, slots.as<unsigned long long>(), order.as<int64_t>(), ng, bit,
                                                     ddof, out, keys_written ? 0 : 1);
            SDC_CHECK_LAUNCH();
            keys_written = true;
        }
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
    for (int i = 0; i < ncols; ++i) res->col_is_f64[i] = c.is_f64[i];
    cudaGetDevice(&res->device);
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
    const int ai = agg_index(agg);
    SDC_ARG_CHECK(ai >= 0 && (r->agg_mask & agg), "aggregate was not requested");
    SDC_ARG_CHECK(col >= 0 && col < r->ncols, "column index");
    if (r->ngroups == 0) return SDCB200_OK;
    SDC_CUDA_TRY(cudaMemcpyAsync(dst, r->vals[ai][col], (size_t)r->ngroups * 8,
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
