// rolling_moments.cu -- Series.rolling(window, min_periods).skew() / .kurt() / .cov(other) / .corr(other).
//
// Replaces (citations relative to /root/reference):
//   skew  sdc/datatypes/hpat_pandas_series_rolling_functions.py:408-431 (put / pop), :548-556 (result),
//         sdc/functions/statistics.py:33-41 (formula)
//   kurt  :303-331, :520-536
//   cov   :261-300 (align_finiteness = True in both overloads, :995-1012), :509-517, ddof :559-565
//   corr  :207-236, :487-506; _almost_equal = |x - y| <= eps (sdc/datatypes/common_functions.py:584-600)
//   the two-series driver (different lengths) :824-879, :925-992
//
// The reference slides sums of powers / cross products along each chunk (add the entering row, subtract the
// leaving one).  Here a CTA stages a tile of rows (plus the win-1 rows in front) in shared memory; thread t owns
// R consecutive outputs: it sums the window of its first output from scratch and slides over the remaining
// R - 1, so the drift of the subtract-what-you-added scheme is bounded by R rows instead of a whole chunk.  R is
// odd: 64-bit shared accesses at a stride of R rows are bank-conflict-free.  A row takes part only when it is
// finite in BOTH series; rows outside [0, min(len_x, len_y)) are staged as NaN, which also gives the reference's
// tail (outputs past min(len) + window are NaN).  Results leave through a shared tile with coalesced stores.
// HBM traffic = every input row once per tile (+ the halo) and every output once: 16 B / row (24 B with `other`).
#include <cstdint>
#include <cuda_runtime.h>

#include "common.cuh"

namespace sdcb200 {
namespace {

enum { OP_SKEW = 7, OP_KURT = 8, OP_COV = 9, OP_CORR = 10 };
constexpr int kMThreads = 256;

struct MomArgs {
    const double* x;
    const double* y;        // == x when `other` is absent
    double* out;
    int64_t n_out;          // max(len_x, len_y)
    int64_t n_min;          // min(len_x, len_y): only these rows enter a window
    int win, minp, ddof;
};

struct MomState {
    int n;
    double s0, s1, s2, s3, s4;
};

template <int OP>
__device__ __forceinline__ void mom_update(MomState& st, double x, double y, double sign, int dn) {
    if (!(is_finite_f64(x) && is_finite_f64(y))) return;
    st.n += dn;
    if (OP == OP_SKEW || OP == OP_KURT) {
        const double x2 = x * x;
        st.s0 += sign * x;
        st.s1 += sign * x2;
        st.s2 += sign * (x2 * x);
        if (OP == OP_KURT) st.s3 += sign * (x2 * x * x);
    } else {
        st.s0 += sign * x;
        st.s1 += sign * y;
        st.s2 += sign * (x * y);
        if (OP == OP_CORR) {
            st.s3 += sign * (x * x);
            st.s4 += sign * (y * y);
        }
    }
}

// Everything in the result formulas that depends on the row count alone, cached while the count stays the same
// (consecutive windows almost always hold the same number of finite rows): the divisions by n become
// multiplications by 1/n, which leaves one division (and one square root) per output.
struct MomConsts {
    int n;              // the count the constants belong to (-1: none yet)
    double rn;          // 1 / n
    double c0, c1, c2;  // skew: sqrt((n-1) n) / (n-2);  kurt: 1 / ((n-2)(n-3)), n^2 - 1, 3 (n-1)^2;  cov: n / (n - ddof)
};

template <int OP>
__device__ __forceinline__ void mom_consts_for(MomConsts& c, int ni, int ddof) {
    if (ni == c.n) return;
    c.n = ni;
    const double n = (double)ni;
    c.rn = ni != 0 ? 1.0 / n : 0.0;
    if (OP == OP_SKEW) c.c0 = ni > 2 ? sqrt((n - 1.0) * n) / (n - 2.0) : 0.0;
    if (OP == OP_KURT) {
        c.c0 = ni > 3 ? 1.0 / (n - 2.0) / (n - 3.0) : 0.0;
        c.c1 = n * n - 1.0;
        c.c2 = 3.0 * (n - 1.0) * (n - 1.0);
    }
    if (OP == OP_COV) c.c0 = (ni - ddof) >= 1 ? n / (double)(ni - ddof) : 0.0;
}

// skew_result_or_nan / kurt_result_or_nan / cov_result_or_nan / corr_result_or_nan (reference :487-556)
template <int OP>
__device__ __forceinline__ double mom_emit(const MomState& st, int minp, int ddof, MomConsts& c) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const int ni = st.n;
    if (OP == OP_SKEW) {
        if (ni < (minp > 3 ? minp : 3)) return nan;
        mom_consts_for<OP>(c, ni, ddof);
        const double a = st.s0 * c.rn;                                  // mean
        const double m2 = (st.s1 - st.s0 * a) * c.rn;
        const double m3 = (st.s2 - 3.0 * a * st.s1 + 2.0 * a * a * st.s0) * c.rn;
        if (!(m2 > 0.0)) return m2 == 0.0 ? nan : m3 / (m2 * sqrt(m2));     // m2 < 0 (rounding): pow gives NaN
        return c.c0 * m3 / (m2 * sqrt(m2));
    }
    if (OP == OP_KURT) {
        if (ni < (minp > 4 ? minp : 4)) return nan;
        mom_consts_for<OP>(c, ni, ddof);
        const double a = st.s0 * c.rn;
        const double m2 = (st.s1 - st.s0 * a) * c.rn;
        const double m4 = (st.s3 - 4.0 * a * st.s2 + 6.0 * a * a * st.s1 - 3.0 * a * a * a * st.s0) * c.rn;
        if (!(m2 > 0.0)) return m2 == 0.0 ? 0.0 : m4 / (m2 * m2);
        return c.c0 * (c.c1 * m4 / (m2 * m2) - c.c2);
    }
    if (ni < (minp > 1 ? minp : 1)) return nan;
    mom_consts_for<OP>(c, ni, ddof);
    if (OP == OP_COV) {
        if (ni - ddof < 1) return nan;
        return (st.s2 - st.s0 * st.s1 * c.rn) * c.rn * c.c0;
    }
    const double eps = 2.220446049250313e-16;
    const double var_x = st.s3 - st.s0 * st.s0 * c.rn;
    if (fabs(var_x) <= eps) return nan;
    const double var_y = st.s4 - st.s1 * st.s1 * c.rn;
    if (fabs(var_y) <= eps) return nan;
    const double cov_xy = st.s2 - st.s0 * st.s1 * c.rn;
    return cov_xy / sqrt(var_x * var_y);
}

// tile = kMThreads * R outputs; staged rows = hp + tile, hp = win - 1 (win >= 1)
template <int OP, int R>
__global__ void __launch_bounds__(kMThreads) rolling_moments_kernel(MomArgs a) {
    constexpr bool kTwo = (OP == OP_COV || OP == OP_CORR);
    constexpr int TO = kMThreads * R;
    extern __shared__ __align__(16) double msm[];
    const int hp = a.win - 1;
    double* X = msm;                                   // [hp + TO]
    double* Y = kTwo ? X + hp + TO : X;                // [hp + TO]
    double* O = (kTwo ? Y : X) + hp + TO;              // [TO]
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const int tid = threadIdx.x;
    const bool same = a.x == a.y;
    for (int64_t tile0 = (int64_t)blockIdx.x * TO; tile0 < a.n_out; tile0 += (int64_t)gridDim.x * TO) {
        const int64_t r0 = tile0 - hp;                 // global row of X[0]
        for (int j = tid; j < hp + TO; j += kMThreads) {
            const int64_t r = r0 + j;
            const bool in = r >= 0 && r < a.n_min;
            X[j] = in ? __ldg(a.x + r) : nan;
            if (kTwo) Y[j] = in ? (same ? X[j] : __ldg(a.y + r)) : nan;
        }
        __syncthreads();
        const int o0 = tid * R;
        if (tile0 + o0 < a.n_out) {
            MomState st = {0, 0.0, 0.0, 0.0, 0.0, 0.0};
            MomConsts mc;
            mc.n = -1;
            auto row = [&](int j, double sign, int dn) {
                const double xv = X[j];
                mom_update<OP>(st, xv, kTwo ? Y[j] : xv, sign, dn);
            };
            for (int j = o0; j <= o0 + hp; ++j) row(j, 1.0, 1);
            O[o0] = mom_emit<OP>(st, a.minp, a.ddof, mc);
#pragma unroll 4
            for (int k = 1; k < R; ++k) {
                row(o0 + hp + k, 1.0, 1);
                row(o0 + k - 1, -1.0, -1);
                O[o0 + k] = mom_emit<OP>(st, a.minp, a.ddof, mc);
            }
        }
        __syncthreads();
        int64_t rows = a.n_out - tile0;
        if (rows > TO) rows = TO;
        for (int j = tid; j < rows; j += kMThreads) a.out[tile0 + j] = O[j];
        __syncthreads();
    }
}

__global__ void fill_nan_kernel(double* out, int64_t n) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = nan;
}

template <int OP, int R>
int32_t launch_moments(const MomArgs& a, cudaStream_t s) {
    constexpr bool kTwo = (OP == OP_COV || OP == OP_CORR);
    constexpr int TO = kMThreads * R;
    auto kern = rolling_moments_kernel<OP, R>;
    const size_t smem = ((size_t)(a.win - 1 + TO) * (kTwo ? 2 : 1) + TO) * 8;
    SDC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SDC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kMThreads, smem));
    if (occ < 1) occ = 1;
    const int64_t ntiles = ceil_div(a.n_out, (int64_t)TO);
    int64_t grid = (int64_t)sm_count() * occ;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, kMThreads, smem, s>>>(a);
    SDC_CHECK_LAUNCH();
    return SDCB200_OK;
}

template <int OP>
int32_t dispatch_moments(const MomArgs& a, cudaStream_t s) {
    // long runs amortise the from-scratch window sum; short ones keep more CTAs per SM when the window is small or
    // two series are staged (R = 31 with two series needs 192 KB: one CTA per SM, 2.3 ms per 100 M rows for cov)
    constexpr bool kTwo = (OP == OP_COV || OP == OP_CORR);
    if (a.win >= (kTwo ? 512 : 48)) return launch_moments<OP, 31>(a, s);
    return launch_moments<OP, 15>(a, s);
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" int32_t sdcb200_rolling_moments(int32_t op, const double* x, int64_t n_x, const double* y, int64_t n_y,
                                           double* out, int64_t win, int64_t minp, int64_t ddof, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    SDC_ARG_CHECK(op >= OP_SKEW && op <= OP_CORR, "op must be 7 (skew), 8 (kurt), 9 (cov) or 10 (corr)");
    SDC_ARG_CHECK(n_x >= 0 && n_y >= 0 && win >= 0 && minp >= 0 && minp <= win, "n/win/minp");
    if (y == nullptr) { y = x; n_y = n_x; }
    SDC_ARG_CHECK((op == OP_COV || op == OP_CORR) || (y == x && n_y == n_x), "skew / kurt take one series");
    MomArgs a;
    a.x = x; a.y = y; a.out = out;
    a.n_out = n_x > n_y ? n_x : n_y;
    a.n_min = n_x < n_y ? n_x : n_y;
    if (a.n_out == 0) return SDCB200_OK;
    SDC_ARG_CHECK(out != nullptr && (a.n_min == 0 || (x != nullptr && y != nullptr)), "null column");
    if (win == 0 || a.n_min == 0) {                    // every window is empty: NaN for all four reductions
        fill_nan_kernel<<<sm_count() * 4, 256, 0, s>>>(out, a.n_out);
        SDC_CHECK_LAUNCH();
        return SDCB200_OK;
    }
    if (win > 4097) {
        set_error("rolling moments: window %lld exceeds the tiled limit 4097", (long long)win);
        return SDCB200_ERR_ARG;
    }
    a.win = (int)win; a.minp = (int)minp;
    a.ddof = (int)(ddof > (1 << 30) ? (1 << 30) : (ddof < -(1 << 30) ? -(1 << 30) : ddof));
    switch (op) {
        case OP_SKEW: return dispatch_moments<OP_SKEW>(a, s);
        case OP_KURT: return dispatch_moments<OP_KURT>(a, s);
        case OP_COV: return dispatch_moments<OP_COV>(a, s);
        default: return dispatch_moments<OP_CORR>(a, s);
    }
}
