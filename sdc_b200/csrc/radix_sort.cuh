// radix_sort.cuh -- internal interface of the onesweep LSD radix sort (sort.cu).
#pragma once
#include "common.cuh"

namespace sdcb200 {

// Stable LSD radix sort of `n` keys of dtype code `dtype` (SDCB200_I8..F64) held as raw bits.
//   src             : read-only input keys;  bufA / bufB : ping-pong outputs (n elements each;
//                     bufB may alias src for an in-place sort)
//   pay_bytes       : 0 (keys only), 4 or 8 bytes of payload per row; psrc / pA / pB likewise
//   iota            : the payload entering the first pass is the row number (psrc unused)
//   desc            : order by `greater`; NaN keys stay last in both directions
//   *rkeys / *rpay  : whichever buffer holds the sorted rows afterwards (src when no pass ran)
// Order = the reference comparators (sdc/native/utils.cpp:160-182): NaN last, -0.0 == +0.0,
// ties keep their input order.  Synchronises the stream once (digit-histogram read-back).
int32_t radix_sort_core(int dtype, const void* src, void* bufA, void* bufB, int pay_bytes, const void* psrc, void* pA,
                        void* pB, int64_t n, bool iota, bool desc, const void** rkeys, const void** rpay,
                        cudaStream_t s, int64_t* idx64_out = nullptr);
// idx64_out (4-byte payload only): the LAST pass writes no keys and stores the payload as int64 into idx64_out
// (*rkeys is then null, *rpay == idx64_out) -- what an argsort wants.

// One UNORDERED partition pass of 8-byte keys (raw bits; key_f64: canonicalise -0.0 / NaN first) and their payload
// (pay_bytes 0 / 4 / 8; iota: the payload is the row number) into 256 buckets by the TOP byte of
// hash_mix64(canonical key) (common.cuh).  part_begin: device u64[257], bucket b = [part_begin[b], part_begin[b+1]).
// Rows of one bucket arrive in no particular order (a tile claims its range of every bucket with one atomic add on
// the bucket's cursor; inside a tile one shared atomic per row).  Hash tables whose slot index is taken from the top bits of the same
// hash are then touched one contiguous 1/256 slice at a time (join / groupby with keys that have no dense range).
int32_t hash_partition_pass(const void* keys, bool key_f64, void* keys_out, int pay_bytes, const void* pay, void* pay_out,
                            int64_t n, bool iota, unsigned long long* part_begin, cudaStream_t s);

// The same pass over a DENSE key range: bucket = (key - lo) >> shift (the caller picks shift so that at most 256 buckets
// cover the range).  Rows of one bucket touch one contiguous slice of a direct-addressed table indexed by key - lo.
int32_t range_partition_pass(const void* keys, unsigned long long lo, int shift, void* keys_out, int pay_bytes, const void* pay,
                             void* pay_out, int64_t n, bool iota, unsigned long long* part_begin, cudaStream_t s);

// stable argsort of a device column -> int64 row numbers
int32_t argsort_device(int dtype, const void* keys, int64_t n, bool ascending, int64_t* index_out, cudaStream_t s);

int dtype_size(int dtype);

}  // namespace sdcb200
