// csv.cu -- CSV text in device memory -> int64 / float64 columns in HBM.
//
// Replaces, for numeric tables, the reference's CSV ingest: pd.read_csv is lowered to pyarrow.csv.read_csv
// (sdc/io/csv_ext.py:93-110, options :174-272; native side sdc/native/arrow_reader.cpp) on the host's cores, after
// which every column is boxed into NumPy and -- with a GPU backend -- copied to the device.  Here the file's bytes
// are copied to HBM once and decoded there:
//   1. csv_mark_kernel   counts the row starts of every 4 KB block (a row starts after a '\n' unless the line is
//                        empty: Arrow's ignore_empty_lines), a scan turns the counts into offsets, and the same
//                        kernel's second mode writes row_start[] -- 8 bytes per row, the only index kept;
//   2. csv_rows_kernel   one thread per row, 256 rows per CTA: the CTA's rows are one contiguous byte range of the
//                        file, copied to shared memory with 16-byte loads (coalesced, each byte of the file is read
//                        from HBM once); a thread then walks its row field by field and converts with num_parse.cuh
//                        (int64 with overflow check, float64 correctly rounded: Eisel-Lemire).  Column c of row r is
//                        stored at out[c][r]: consecutive threads, consecutive addresses.
//      mode CLASSIFY ORs the kind of every field (int / float / null / other) into kinds[c] -- type inference as
//      Arrow does it (int64 if every field is an int, else double, else string);
//      mode PARSE writes the columns; the first offending field (row, column, reason) is reported.
// Quoted fields are not handled (a '"' is reported as an error): string columns go through Arrow, whose buffers are
// the str_arr layout already (sdc_b200/str_arr.py from_arrow).
#include "num_parse.cuh"
#include "scan.cuh"

namespace sdcb200 {
namespace {

__device__ const uint64_t kPow5Dev[] = {
#include "pow5_table.inc"
};

constexpr int kMarkThreads = 256;
constexpr int kMarkBytes = 16;                               // bytes per thread
constexpr int kMarkBlock = kMarkThreads * kMarkBytes;        // 4 KB per CTA

__device__ __forceinline__ bool row_starts_at(const uint8_t* __restrict__ d, int64_t i, int64_t n, uint8_t prev) {
    if (prev != '\n') return false;
    const uint8_t c = d[i];
    if (c == '\n') return false;                             // empty line
    if (c == '\r' && (i + 1 == n || d[i + 1] == '\n')) return false;
    return true;
}

// mode 0: counts[block] = row starts in the block; mode 1: row_start[offsets[block] + k] = position of the k-th
template <int MODE>
__global__ void __launch_bounds__(kMarkThreads) csv_mark_kernel(const uint8_t* __restrict__ d, int64_t n,
                                                                unsigned long long* __restrict__ counts,
                                                                const unsigned long long* __restrict__ offsets,
                                                                unsigned long long* __restrict__ row_start) {
    __shared__ unsigned wsum[kMarkThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b0 = (int64_t)blockIdx.x * kMarkBlock + (int64_t)tid * kMarkBytes;
    unsigned mask = 0;                                       // bit k: a row starts at b0 + k
    if (b0 < n) {
        uint8_t prev = b0 == 0 ? (uint8_t)'\n' : d[b0 - 1];
        const int m = (int)(n - b0 < kMarkBytes ? n - b0 : kMarkBytes);
        for (int k = 0; k < m; ++k) {
            if (row_starts_at(d, b0 + k, n, prev)) mask |= 1u << k;
            prev = d[b0 + k];
        }
    }
    const unsigned c = (unsigned)__popc(mask);
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kMarkThreads / 32; ++w) {
        if (w < warp) base += wsum[w];
        tot += wsum[w];
    }
    if (MODE == 0) {
        if (tid == 0) counts[blockIdx.x] = tot;
    } else {
        unsigned long long j = offsets[blockIdx.x] + base + incl - c;
        while (mask) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            row_start[j++] = (unsigned long long)(b0 + k);
        }
    }
}

struct CountLoad {
    const unsigned long long* c;
    __device__ unsigned long long operator()(int64_t i) const { return c[i]; }
};
struct OffsetStore64 {
    unsigned long long* o;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long) const { o[i] = excl; }
};

enum { CSV_SKIP = 0, CSV_I64 = 1, CSV_F64 = 2 };
enum { KIND_INT = 1, KIND_FLOAT = 2, KIND_NULL = 4, KIND_OTHER = 8 };
enum { ROW_BAD_VALUE = NP_BAD, ROW_RANGE = NP_RANGE, ROW_AMBIG = NP_AMBIG, ROW_NULL_INT = 5, ROW_FEW = 6, ROW_MANY = 7, ROW_QUOTE = 8 };

constexpr int kRowThreads = 256;                             // most rows per CTA
constexpr int kRowSmem = 96 * 1024;                          // most staged bytes per CTA; a longer span is read from global memory

// error word: row << 20 | column << 4 | reason -- atomicMin keeps the first offending row
__device__ __forceinline__ void report(unsigned long long* err, int64_t row, int col, int why) {
    atomicMin(err, ((unsigned long long)row << 20) | ((unsigned long long)(col & 0xffff) << 4) | (unsigned long long)why);
}

template <bool CLASSIFY>
__global__ void __launch_bounds__(kRowThreads) csv_rows_kernel(const uint8_t* __restrict__ d, int64_t nbytes,
                                                               const unsigned long long* __restrict__ row_start, int64_t nrows,
                                                               int ncols, uint8_t delim, const int* __restrict__ types,
                                                               void* const* __restrict__ out, unsigned* __restrict__ kinds,
                                                               unsigned long long* __restrict__ err, int smem_bytes) {
    extern __shared__ __align__(16) uint8_t stage[];
    const int tid = threadIdx.x;
    const int nthr = (int)blockDim.x;                         // rows per CTA: 256 for narrow rows, fewer for wide ones
    const int64_t r0 = (int64_t)blockIdx.x * nthr;
    const int64_t r1 = r0 + nthr < nrows ? r0 + nthr : nrows;
    // the CTA's rows are the bytes [lo, hi) of the file; staged from the 16-byte boundary below lo
    const int64_t lo = (int64_t)row_start[r0];
    const int64_t hi = r1 < nrows ? (int64_t)row_start[r1] : nbytes;
    // (the 16-byte granule is taken on the ADDRESS: `d` is usually the buffer's start plus the header's length; a granule
    // that holds a byte of the data lies in the same allocation page, so reading it whole is safe)
    const int64_t alo = lo - (int64_t)(reinterpret_cast<uintptr_t>(d + lo) & 15);
    const bool staged = hi - alo <= (int64_t)smem_bytes;
    if (staged) {
        const int64_t nvec = (hi - alo) >> 4;                 // whole 16-byte words, then the tail byte by byte
        const uint4* src = reinterpret_cast<const uint4*>(d + alo);
        uint4* dst = reinterpret_cast<uint4*>(stage);
        for (int64_t v = tid; v < nvec; v += nthr) dst[v] = src[v];
        for (int64_t b = alo + (nvec << 4) + tid; b < hi; b += nthr) stage[b - alo] = d[b];
    }
    __syncthreads();
    const int64_t r = r0 + tid;
    if (r >= nrows) return;
    const uint8_t* base = staged ? stage - alo : d;           // base[i] = byte i of the file
    int64_t p = (int64_t)row_start[r];
    for (int c = 0;; ++c) {
        const int64_t fb = p;
        uint8_t ch = 0;
        bool quote = false;
        while (p < nbytes) {
            ch = base[p];
            if (ch == delim || ch == '\n') break;
            quote |= ch == '"';
            ++p;
        }
        const bool last = p >= nbytes || ch == '\n';
        int64_t fe = p;
        if (last && fe > fb && base[fe - 1] == '\r') --fe;     // CRLF line ends
        if (c >= ncols) { report(err, r, c, ROW_MANY); return; }
        if (quote && !CLASSIFY) { report(err, r, c, ROW_QUOTE); return; }
        const uint8_t* s = base + fb;
        const int len = (int)(fe - fb);
        if (CLASSIFY) {
            int64_t iv;
            double fv;
            int k;
            const int si = parse_int64_field(s, len, &iv);
            if (quote) k = KIND_OTHER;                         // a quoted field: not a column this reader decodes
            else if (si == NP_OK) k = KIND_INT;
            else if (si == NP_EMPTY) k = KIND_NULL;
            else {
                const int sf = parse_f64_field(s, len, kPow5Dev, &fv);
                k = (sf == NP_OK || sf == NP_AMBIG) ? KIND_FLOAT : KIND_OTHER;
            }
            if (!(kinds[c] & (unsigned)k)) atomicOr(&kinds[c], (unsigned)k);
        } else {
            const int t = types[c];
            if (t == CSV_I64) {
                int64_t v = 0;
                const int st = parse_int64_field(s, len, &v);
                if (st != NP_OK) { report(err, r, c, st == NP_EMPTY ? ROW_NULL_INT : st); return; }
                reinterpret_cast<int64_t*>(out[c])[r] = v;
            } else if (t == CSV_F64) {
                double v = 0.0;
                const int st = parse_f64_field(s, len, kPow5Dev, &v);
                if (st != NP_OK && st != NP_EMPTY) { report(err, r, c, st); return; }
                reinterpret_cast<double*>(out[c])[r] = v;
            }
        }
        if (last) {
            if (c + 1 < ncols) report(err, r, c + 1, ROW_FEW);
            return;
        }
        ++p;                                                   // past the delimiter
    }
}

struct CsvIndex {
    const uint8_t* data;
    int64_t nbytes;
    unsigned long long* row_start;      // device, nrows entries
    int64_t nrows;
    cudaStream_t stream;
};

const char* why_text(int why) {
    switch (why) {
        case ROW_BAD_VALUE: return "not a number";
        case ROW_RANGE: return "out of the int64 range";
        case ROW_AMBIG: return "more than 19 significant digits and too close to a rounding boundary for the device reader (use engine='arrow')";
        case ROW_NULL_INT: return "missing value in an int64 column (read it as float64)";
        case ROW_FEW: return "too few fields in the row";
        case ROW_MANY: return "too many fields in the row";
        case ROW_QUOTE: return "quoted fields are not supported by the device reader (use engine='arrow')";
    }
    return "?";
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int32_t sdcb200_csv_open(const uint8_t* data, int64_t nbytes, int64_t* handle_out, int64_t* nrows_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(nbytes >= 0 && handle_out != nullptr && nrows_out != nullptr, "args");
    *handle_out = 0;
    *nrows_out = 0;
    SDC_ARG_CHECK(nbytes == 0 || data != nullptr, "null data");
    CsvIndex* ix = new CsvIndex{data, nbytes, nullptr, 0, s};
    if (nbytes > 0) {
        const int64_t nb = ceil_div(nbytes, kMarkBlock);
        Scratch counts, offsets, tsum;
        int32_t rc = counts.alloc((size_t)nb * 8, s);
        if (rc == SDCB200_OK) rc = offsets.alloc((size_t)nb * 8, s);
        if (rc == SDCB200_OK) rc = tsum.alloc((size_t)(ceil_div(nb, kScanTile) + 1) * 8, s);
        if (rc != SDCB200_OK) { delete ix; return rc; }
        auto fail = [&](int32_t code) { if (ix->row_start) dev_free(ix->row_start, s); delete ix; return code; };
        csv_mark_kernel<0><<<(unsigned)nb, kMarkThreads, 0, s>>>(data, nbytes, counts.as<unsigned long long>(), nullptr, nullptr);
        if (cudaGetLastError() != cudaSuccess) { set_error("csv: launch failed"); return fail(SDCB200_ERR_CUDA); }
        rc = device_exclusive_scan(CountLoad{counts.as<unsigned long long>()}, OffsetStore64{offsets.as<unsigned long long>()}, nb,
                                   tsum.as<unsigned long long>(), s);
        if (rc != SDCB200_OK) return fail(rc);
        unsigned long long total = 0;
        if (cudaMemcpyAsync(&total, tsum.as<unsigned long long>() + ceil_div(nb, kScanTile), 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) { set_error("csv: row count read-back failed"); return fail(SDCB200_ERR_CUDA); }
        ix->nrows = (int64_t)total;
        rc = dev_alloc(reinterpret_cast<void**>(&ix->row_start), (size_t)(total ? total : 1) * 8, s);
        if (rc != SDCB200_OK) return fail(rc);
        if (total) {
            csv_mark_kernel<1><<<(unsigned)nb, kMarkThreads, 0, s>>>(data, nbytes, nullptr, offsets.as<unsigned long long>(), ix->row_start);
            if (cudaGetLastError() != cudaSuccess) { set_error("csv: launch failed"); return fail(SDCB200_ERR_CUDA); }
        }
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("csv: index kernels failed"); return fail(SDCB200_ERR_CUDA); }
    }
    *handle_out = (int64_t)reinterpret_cast<uintptr_t>(ix);
    *nrows_out = ix->nrows;
    return SDCB200_OK;
}

static int32_t csv_run(CsvIndex* ix, bool classify, int32_t ncols, uint8_t delim, const int32_t* types, void* const* out_cols,
                       uint32_t* kinds_out) {
    cudaStream_t s = ix->stream;
    auto kern = classify ? csv_rows_kernel<true> : csv_rows_kernel<false>;
    SDC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmem));
    // rows per CTA and staging bytes from the average row length (x 1.25 for the spread): narrow rows 256 per CTA,
    // wide ones (census-shaped: 24 float columns, ~450 bytes) fewer, so that the CTA's byte range still fits
    const double avg = ix->nrows > 0 ? (double)ix->nbytes / (double)ix->nrows : 1.0;
    int threads = kRowThreads;
    while (threads > 32 && avg * 1.25 * threads + 64 > (double)kRowSmem) threads >>= 1;
    int smem = (int)(avg * 1.25 * threads) + 64;
    smem = smem < 8192 ? 8192 : (smem > kRowSmem ? kRowSmem : smem);
    smem = (smem + 15) & ~15;
    Scratch dtypes, dout, dkinds, derr;
    SDC_TRY(derr.alloc(8, s));
    const unsigned long long none = ~0ull;
    SDC_CUDA_TRY(cudaMemcpyAsync(derr.p, &none, 8, cudaMemcpyHostToDevice, s));
    if (classify) {
        SDC_TRY(dkinds.alloc((size_t)ncols * 4, s));
        SDC_CUDA_TRY(cudaMemsetAsync(dkinds.p, 0, (size_t)ncols * 4, s));
    } else {
        SDC_TRY(dtypes.alloc((size_t)ncols * 4, s));
        SDC_TRY(dout.alloc((size_t)ncols * 8, s));
        SDC_CUDA_TRY(cudaMemcpyAsync(dtypes.p, types, (size_t)ncols * 4, cudaMemcpyHostToDevice, s));
        SDC_CUDA_TRY(cudaMemcpyAsync(dout.p, out_cols, (size_t)ncols * 8, cudaMemcpyHostToDevice, s));
    }
    if (ix->nrows > 0) {
        const int64_t grid = ceil_div(ix->nrows, threads);
        kern<<<(unsigned)grid, threads, smem, s>>>(ix->data, ix->nbytes, ix->row_start, ix->nrows, ncols, delim,
                                                   dtypes.as<int>(), reinterpret_cast<void* const*>(dout.p),
                                                   dkinds.as<unsigned>(), derr.as<unsigned long long>(), smem);
        SDC_CHECK_LAUNCH();
    }
    unsigned long long err = 0;
    SDC_CUDA_TRY(cudaMemcpyAsync(&err, derr.p, 8, cudaMemcpyDeviceToHost, s));
    if (classify) SDC_CUDA_TRY(cudaMemcpyAsync(kinds_out, dkinds.p, (size_t)ncols * 4, cudaMemcpyDeviceToHost, s));
    SDC_CUDA_TRY(cudaStreamSynchronize(s));
    if (err != none) {
        set_error("csv: data row %lld, column %d: %s", (long long)(err >> 20), (int)((err >> 4) & 0xffff), why_text((int)(err & 15)));
        return SDCB200_ERR_ARG;
    }
    return SDCB200_OK;
}

int32_t sdcb200_csv_classify(int64_t handle, int32_t ncols, int32_t delimiter, uint32_t* kinds_out) {
    CsvIndex* ix = reinterpret_cast<CsvIndex*>((uintptr_t)handle);
    SDC_ARG_CHECK(ix != nullptr && ncols > 0 && ncols < 65536 && kinds_out != nullptr, "args");
    return csv_run(ix, true, ncols, (uint8_t)delimiter, nullptr, nullptr, kinds_out);
}

int32_t sdcb200_csv_parse(int64_t handle, int32_t ncols, int32_t delimiter, const int32_t* col_types, void* const* out_cols) {
    CsvIndex* ix = reinterpret_cast<CsvIndex*>((uintptr_t)handle);
    SDC_ARG_CHECK(ix != nullptr && ncols > 0 && ncols < 65536 && col_types != nullptr && out_cols != nullptr, "args");
    for (int c = 0; c < ncols; ++c) {
        SDC_ARG_CHECK(col_types[c] >= CSV_SKIP && col_types[c] <= CSV_F64, "column type (0 skip, 1 int64, 2 float64)");
        SDC_ARG_CHECK(col_types[c] == CSV_SKIP || ix->nrows == 0 || out_cols[c] != nullptr, "null output column");
    }
    return csv_run(ix, false, ncols, (uint8_t)delimiter, col_types, out_cols, nullptr);
}

int32_t sdcb200_csv_close(int64_t handle) {
    CsvIndex* ix = reinterpret_cast<CsvIndex*>((uintptr_t)handle);
    if (!ix) return SDCB200_OK;
    if (ix->row_start) dev_free(ix->row_start, ix->stream);
    delete ix;
    return SDCB200_OK;
}

}  // extern "C"
