// num_parse.cuh -- one CSV field -> int64 / float64, for device code and (the same source) a host test harness.
//
// The reference reads CSV through Arrow's reader (sdc/io/csv_ext.py:93-110 -> pyarrow.csv.read_csv; the native side is
// sdc/native/arrow_reader.cpp), whose number conversion is correctly rounded (Arrow vendors fast_float).  Bit-identical
// columns therefore need a correctly rounded decimal -> binary64 conversion here too: the Eisel-Lemire algorithm
// (Lemire, "Number parsing at a gigabyte per second", 2021; Mushtak & Lemire, "Fast number parsing without fallback",
// 2023 -- no fallback is needed for up to 19 significant digits) over a 128-bit table of powers of five
// (pow5_table.inc, produced by tools/gen_pow5_table.py).  Longer digit strings are truncated to 19 digits and
// converted as w and w + 1: when both round to the same double that is the answer, otherwise (about one such field in
// a thousand) the full digit string is compared with the midpoint of the two candidates in exact big-integer
// arithmetic (round_long_decimal below) -- what Arrow's fast_float does with its big-decimal fallback.
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

// ---- more than 19 significant digits whose 19-digit bracket [w, w + 1] straddles a rounding boundary ----
// The two candidates are adjacent doubles b and b + 1 ulp; the answer depends on which side of their midpoint the
// full decimal value lies.  Both are compared exactly as big integers (32-bit limbs):
//     D * 10^q   against   (2 m + 1) * 2^(e - 1)          (b = m * 2^e, D = all significant digits)
// with the powers of five multiplied onto the side where their exponent is positive and the powers of two applied as
// one shift.  Up to kBigDigits digits take part; non-zero digits beyond them only matter for a tie (they break it
// upwards).  A rare path (about one in a thousand fields with more than 19 digits): kept out of line.
constexpr int kBigLimbs = 192;               // 6144 bits: 800 digits (2658 bits) * 5^308 (716 bits) shifted by up to 2^1140
constexpr int kBigDigits = 800;

struct BigNum {
    uint32_t v[kBigLimbs];
    int n;                                   // limbs in use (v[n - 1] != 0), 0 = zero
};
SDC_HD void big_set(BigNum& a, uint64_t x) {
    a.n = 0;
    while (x) { a.v[a.n++] = (uint32_t)x; x >>= 32; }
}
SDC_HD void big_mul_add(BigNum& a, uint32_t mul, uint32_t add) {
    uint64_t carry = add;
    for (int i = 0; i < a.n; ++i) {
        const uint64_t t = (uint64_t)a.v[i] * mul + carry;
        a.v[i] = (uint32_t)t;
        carry = t >> 32;
    }
    if (carry && a.n < kBigLimbs) a.v[a.n++] = (uint32_t)carry;
}
SDC_HD void big_mul_pow5(BigNum& a, int64_t p) {
    while (p >= 13) { big_mul_add(a, 1220703125u, 0u); p -= 13; }          // 5^13
    uint32_t m = 1;
    while (p-- > 0) m *= 5u;
    if (m > 1) big_mul_add(a, m, 0u);
}
SDC_HD void big_shl(BigNum& a, int64_t bits) {
    if (a.n == 0 || bits <= 0) return;
    const int limbs = (int)(bits >> 5), r = (int)(bits & 31);
    if (a.n + limbs + 1 > kBigLimbs) return;                                // cannot happen within the documented bounds
    for (int i = a.n - 1; i >= 0; --i) {
        const uint64_t t = (uint64_t)a.v[i] << r;
        if (i == a.n - 1) a.v[i + limbs + 1] = (uint32_t)(t >> 32);
        else a.v[i + limbs + 1] |= (uint32_t)(t >> 32);
        a.v[i + limbs] = (uint32_t)t;
    }
    for (int i = 0; i < limbs; ++i) a.v[i] = 0;
    a.n += limbs + 1;
    while (a.n > 0 && a.v[a.n - 1] == 0) --a.n;
}
SDC_HD int big_cmp(const BigNum& a, const BigNum& b) {
    if (a.n != b.n) return a.n < b.n ? -1 : 1;
    for (int i = a.n - 1; i >= 0; --i)
        if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1;
    return 0;
}

// s[0, len): the digits part of the field ("123.456", sign and exponent stripped), exp_explicit the written exponent,
// lower = eisel_lemire(w, ...) of the 19-digit truncation (w <= true value).  -> the correctly rounded bits
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
uint64_t round_long_decimal(const uint8_t* s, int len, int64_t exp_explicit, uint64_t lower) {
    BigNum A, B;
    A.n = 0;
    int nd = 0;
    int64_t q = exp_explicit;
    bool sticky = false, frac = false;
    for (int i = 0; i < len; ++i) {
        if (s[i] == '.') { frac = true; continue; }
        const uint32_t d = (uint32_t)s[i] - '0';
        if (nd < kBigDigits) {
            if (A.n == 0) { if (d) { A.v[0] = d; A.n = 1; } }
            else big_mul_add(A, 10u, d);
            nd += (A.n != 0);
            if (frac) --q;
        } else {
            sticky |= d != 0u;
            if (!frac) ++q;
        }
    }
    // b = m * 2^e; the midpoint to the next double is (2 m + 1) * 2^(e - 1)
    const uint64_t frac_bits = lower & 0x000fffffffffffffull;
    const int64_t bexp = (int64_t)(lower >> 52);
    const uint64_t m = bexp ? (frac_bits | (1ull << 52)) : frac_bits;
    const int64_t e = (bexp ? bexp : 1) - 1075;
    big_set(B, 2ull * m + 1ull);
    int64_t two = q - (e - 1);              // A * 5^q * 2^two  vs  B   (q >= 0)   |   A * 2^two  vs  B * 5^-q   (q < 0)
    if (q >= 0) big_mul_pow5(A, q); else big_mul_pow5(B, -q);
    if (two >= 0) big_shl(A, two); else big_shl(B, -two);
    int c = big_cmp(A, B);
    if (c == 0 && sticky) c = 1;
    if (c > 0) return lower + 1ull;
    if (c < 0) return lower;
    return (lower & 1ull) ? lower + 1ull : lower;          // an exact tie: to even
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
    int64_t exp10 = 0, exp_explicit = 0;
    const int digits_begin = i;
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
    const int digits_end = i;
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
        exp_explicit = eneg ? -e : e;
        exp10 += exp_explicit;
    }
    if (i != len) return NP_BAD;
    if (was_int) *was_int = plain_int && !(signed_ && !neg);              // "+12" is a float to Arrow, not an int
    uint64_t bits = eisel_lemire(w, exp10, pow5);
    if (truncated && eisel_lemire(w + 1ull, exp10, pow5) != bits)
        bits = round_long_decimal(s + digits_begin, digits_end - digits_begin, exp_explicit, bits);
    *out = bits_to_f64(sign | bits);
    return NP_OK;
}

}  // namespace sdcb200
