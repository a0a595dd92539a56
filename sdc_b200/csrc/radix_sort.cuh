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
                        cudaStream_t s);

// stable argsort of a device column -> int64 row numbers
int32_t argsort_device(int dtype, const void* keys, int64_t n, bool ascending, int64_t* index_out, cudaStream_t s);

int dtype_size(int dtype);

}  // namespace sdcb200
