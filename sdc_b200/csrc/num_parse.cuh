// num_parse.cuh -- one CSV field -> int64 / float64, for device code and (the same source) a host test harness.
//
// The reference reads CSV through Arrow's reader (sdc/io/csv_ext.py:93-110 -> pyarrow.csv.read_csv; the native side is
// sdc/native/arrow_reader.cpp), whose number conversion is correctly rounded (Arrow vendors fast_float).  Bit-identical
// columns therefore need a correctly rounded decimal -> binary64 conversion here too: the Eisel-Lemire algorithm
// (Lemire, "Number parsing at a gigabyte per second", 2021; Mushtak & Lemire, "Fast number parsing without fallback",
// 2023 -- no fallback is needed for up to 19 significant digits) over a 128-bit table of powers of five
// (pow5_table.inc, produced by tools/gen_pow5_table.py).  Longer digit strings are truncated to 19 digits and
// converted as w and w + 1: when both round to the same double that is the answer, otherwise the field is reported as
// NP_AMBIG (the caller fails loudly; Arrow resolves those with big-decimal arithmetic).
//
// Grammar = what Arrow accepts for a numeric column (probed against pyarrow 24, tests/test_csv_*):
//   float  : blanks* [+-]? ( digit+ ('.' digit*)? | '.' digit+ ) ([eE] [+-]? digit+)? blanks*
//            | [+-]? ("inf" | "infinity" | "nan") in any case
//            | a null token ("", "NA", "N/A", "NaN", "null", ... Arrow's default null_values)  -> NaN
//   int64  : blanks* '-'? digit+ blanks*   (no '+', no overflow)          | a null token -> NP_EMPTY
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SDC_HD __host__ __device__ __forceinline__
#else
#define SDC_HD inline
#endif

namespace sdcb200 {

enum NumParseStatus { NP_OK = 0, NP_EMPTY = 1, NP_BAD = 2, NP_RANGE = 3, NP_AMBIG = 4 };

constexpr int kPow5Smallest = -342, kPow5Largest = 308;

struct U128 { uint64_t hi, lo; };

SDC_HD U128 mul64x64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return U128{__umul64hi(a, b), a * b};
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    return U128{(uint64_t)(p >> 64), (uint64_t)p};
#endif
}

SDC_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

SDC_HD double bits_to_f64(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    union { uint64_t u; double d; } x;
    x.u = b;
    return x.d;
#endif
}

// w * 10^q (w != 0) -> the nearest binary64, ties to even, as raw bits without the sign.  pow5: the 2 x 651 table.
SDC_HD uint64_t eisel_lemire(uint64_t w, int64_t q, const uint64_t* __restrict__ pow5) {
    if (w == 0 || q < kPow5Smallest) return 0ull;
    if (q > kPow5Largest) return 0x7ff0000000000000ull;
    const int lz = clz64(w);
    w <<= lz;
    // 128-bit product with the truncated power of five; the low table word only when the high product cannot decide
    const uint64_t* e = pow5 + 2 * (q - kPow5Smallest);
    U128 p = mul64x64(w, e[0]);
    if ((p.hi & 0x1ffull) == 0x1ffull) {
        const U128 p2 = mul64x64(w, e[1]);
        p.lo += p2.hi;
        if (p2.hi > p.lo) ++p.hi;
    }
    const int upper = (int)(p.hi >> 63);
    const int shift = upper + 64 - 52 - 3;
    uint64_t m = p.hi >> shift;
    // floor(log2(10^q)) + 63, then the biased binary exponent
    int64_t power2 = ((((int64_t)152170 + 65536) * q) >> 16) + 63 + upper - lz + 1023;
    if (power2 <= 0) {                                  // subnormal (or zero)
        if (-power2 + 1 >= 64) return 0ull;
        m >>= (-power2 + 1);
        m += (m & 1ull);
        m >>= 1;
        return m;                                       // m < 2^52: exponent field 0; m == 2^52: it carries into exponent 1
    }
    // exactly halfway between two doubles: only possible for 5^q that fits 64 bits; round to even
    if (p.lo <= 1ull && q >= -4 && q <= 23 && (m & 3ull) == 1ull) {
        if ((m << shift) == p.hi) m &= ~1ull;
    }
    m += (m & 1ull);
    m >>= 1;
    if (m >= (2ull << 52)) {
        m = 1ull << 52;
        ++power2;
    }
    m &= ~(1ull << 52);
    if (power2 >= 0x7ff) return 0x7ff0000000000000ull;
    return m | ((uint64_t)power2 << 52);
}

SDC_HD bool is_blank(uint8_t c) { return c == ' ' || c == '\t'; }
SDC_HD uint8_t lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }

SDC_HD bool word_is(const uint8_t* s, int len, const char* w, int wl) {          // case-insensitive
    if (len != wl) return false;
    for (int i = 0; i < wl; ++i)
        if (lower(s[i]) != (uint8_t)w[i]) return false;
    return true;
}
SDC_HD bool word_is_exact(const uint8_t* s, int len, const char* w, int wl) {
    if (len != wl) return false;
    for (int i = 0; i < wl; ++i)
        if (s[i] != (uint8_t)w[i]) return false;
    return true;
}

// Arrow's default null spellings (pyarrow.csv.ConvertOptions.null_values), case-sensitive as Arrow matches them
SDC_HD bool is_null_token(const uint8_t* s, int len) {
    if (len == 0) return true;
    if (len > 8) return false;
    return word_is_exact(s, len, "NA", 2) || word_is_exact(s, len, "N/A", 3) || word_is_exact(s, len, "n/a", 3) ||
           word_is_exact(s, len, "NaN", 3) || word_is_exact(s, len, "nan", 3) || word_is_exact(s, len, "-NaN", 4) ||
           word_is_exact(s, len, "-nan", 4) || word_is_exact(s, len, "NULL", 4) || word_is_exact(s, len, "null", 4) ||
           word_is_exact(s, len, "#N/A", 4) || word_is_exact(s, len, "#NA", 3) || word_is_exact(s, len, "#N/A N/A", 8) ||
           word_is_exact(s, len, "1.#IND", 6) || word_is_exact(s, len, "1.#QNAN", 7) || word_is_exact(s, len, "-1.#IND", 7) ||
           word_is_exact(s, len, "-1.#QNAN", 8);
}

SDC_HD void trim_blanks(const uint8_t*& s, int& len) {
    while (len > 0 && is_blank(s[0])) { ++s; --len; }
    while (len > 0 && is_blank(s[len - 1])) --len;
}

SDC_HD int parse_int64_field(const uint8_t* s, int len, int64_t* out) {
    if (is_null_token(s, len)) return NP_EMPTY;
    trim_blanks(s, len);
    if (len == 0) return NP_BAD;
    int i = 0;
    const bool neg = s[0] == '-';
    if (neg) ++i;
    if (i == len) return NP_BAD;
    uint64_t v = 0;
    for (; i < len; ++i) {
        const unsigned d = (unsigned)s[i] - '0';
        if (d > 9u) return NP_BAD;
        if (v > 1844674407370955161ull) return NP_RANGE;                 // v * 10 would wrap
        v = v * 10ull + d;
        if (v < d) return NP_RANGE;
    }
    if (neg) {
        if (v > 9223372036854775808ull) return NP_RANGE;
        *out = (int64_t)(0ull - v);
    } else {
        if (v > 9223372036854775807ull) return NP_RANGE;
        *out = (int64_t)v;
    }
    return NP_OK;
}

// kind of a field for type inference: 1 = int64, 2 = float64 (not an int), 4 = null token, 8 = anything else
SDC_HD int parse_f64_field(const uint8_t* s, int len, const uint64_t* __restrict__ pow5, double* out, bool* was_int = nullptr) {
    const uint64_t kNaN = 0x7ff8000000000000ull;
    if (was_int) *was_int = false;
    if (is_null_token(s, len)) { *out = bits_to_f64(kNaN); return NP_EMPTY; }
    trim_blanks(s, len);
    if (len == 0) return NP_BAD;
    int i = 0;
    bool neg = false, signed_ = false;
    if (s[0] == '-') { neg = true; signed_ = true; ++i; }
    else if (s[0] == '+') { signed_ = true; ++i; }
    if (i == len) return NP_BAD;
    const uint64_t sign = neg ? 0x8000000000000000ull : 0ull;
    if ((unsigned)s[i] - '0' > 9u && s[i] != '.') {
        if (word_is(s + i, len - i, "inf", 3) || word_is(s + i, len - i, "infinity", 8)) { *out = bits_to_f64(sign | 0x7ff0000000000000ull); return NP_OK; }
        if (word_is(s + i, len - i, "nan", 3)) { *out = bits_to_f64(kNaN); return NP_OK; }
        return NP_BAD;
    }
    // up to 19 significant digits into w; further ones only move the exponent (and are remembered when non-zero)
    uint64_t w = 0;
    int nd = 0;                 // significant digits held in w
    int64_t exp10 = 0;
    bool truncated = false, any_digit = false, plain_int = true;
    for (; i < len; ++i) {
        const unsigned d = (unsigned)s[i] - '0';
        if (d > 9u) break;
        any_digit = true;
        if (nd < 19) { w = w * 10ull + d; nd += (w != 0ull); }
        else { ++exp10; truncated |= d != 0u; }
    }
    if (i < len && s[i] == '.') {
        plain_int = false;
        ++i;
        for (; i < len; ++i) {
            const unsigned d = (unsigned)s[i] - '0';
            if (d > 9u) break;
            any_digit = true;
            if (nd < 19) { w = w * 10ull + d; nd += (w != 0ull); --exp10; }
            else truncated |= d != 0u;
        }
    }
    if (!any_digit) return NP_BAD;
    if (i < len && (s[i] == 'e' || s[i] == 'E')) {
        plain_int = false;
        ++i;
        bool eneg = false;
        if (i < len && (s[i] == '-' || s[i] == '+')) { eneg = s[i] == '-'; ++i; }
        if (i == len) return NP_BAD;
        int64_t e = 0;
        for (; i < len; ++i) {
            const unsigned d = (unsigned)s[i] - '0';
            if (d > 9u) return NP_BAD;
            if (e < 100000) e = e * 10 + d;
        }
        exp10 += eneg ? -e : e;
    }
    if (i != len) return NP_BAD;
    if (was_int) *was_int = plain_int && !(signed_ && !neg);              // "+12" is a float to Arrow, not an int
    uint64_t bits = eisel_lemire(w, exp10, pow5);
    if (truncated && eisel_lemire(w + 1ull, exp10, pow5) != bits) return NP_AMBIG;
    *out = bits_to_f64(sign | bits);
    return NP_OK;
}

}  // namespace sdcb200
