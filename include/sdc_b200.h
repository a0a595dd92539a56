/*
 * sdc_b200.h -- C ABI of libsdc_b200.so, the B200 (sm_100a) relational-kernel
 * backend for the data-parallel hot path of IntelPython/sdc.
 *
 * Conventions
 *   - Plain C types only.  `stream` is a cudaStream_t passed as void* (NULL = the
 *     legacy default stream).  Every function returns an int32 status
 *     (SDCB200_OK == 0); `sdcb200_last_error()` gives the message of the last
 *     failure on the calling thread.  (The reference natives return void and have
 *     no error path -- sdc/native/module.cpp:31-92 -- so user-visible argument
 *     errors are still raised by the Python layer before the call.)
 *   - Functions named `sdcb200_<op>` take DEVICE pointers and are asynchronous on
 *     `stream`.  Functions named `sdcb200_host_<op>` take HOST pointers (what the
 *     reference's natives receive from Numba: NumPy buffers), stage through
 *     pinned memory and return after the result is in the caller's buffer.
 *   - The symbols at the end with the reference's own names
 *     (`parallel_stable_argsort_u64f64`, `stable_argsort`, ...) have the exact
 *     signatures the reference binds through ctypes / llvmlite.add_symbol, so
 *     they can replace `sdc.concurrent_sort` / `sdc.hstr_ext` entry points 1:1.
 *
 * Reference citations are relative to /root/reference (IntelPython/sdc).
 */
#ifndef SDC_B200_H
#define SDC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
#define SDCB200_OK            0
#define SDCB200_ERR_ARG       1   /* bad argument (negative length, bad op ...)       */
#define SDCB200_ERR_CUDA      2   /* a CUDA runtime call or kernel launch failed      */
#define SDCB200_ERR_NOMEM     3   /* device / pinned allocation failed                */
#define SDCB200_ERR_CAPACITY  4   /* caller-provided output too small                 */
#define SDCB200_ERR_HANDLE    5   /* unknown / stale handle                           */

const char* sdcb200_last_error(void);
const char* sdcb200_version(void);
/* number of kernel launches issued by this library on the calling process so far
 * (bench.py reports the delta over the timed region as "gpu_launches").        */
int64_t sdcb200_launch_count(void);
int32_t sdcb200_device_info(int32_t* sm_count, int64_t* hbm_bytes, int32_t* cc_major, int32_t* cc_minor);

/* pinned host staging + device scratch used by the host-buffer entry points   */
int32_t sdcb200_host_alloc(void** p, int64_t bytes);          /* cudaHostAlloc  */
int32_t sdcb200_host_free(void* p);
int32_t sdcb200_dev_alloc(void** p, int64_t bytes, void* stream);
int32_t sdcb200_dev_free(void* p, void* stream);
int32_t sdcb200_memcpy_h2d(void* dst, const void* src, int64_t bytes, void* stream);
int32_t sdcb200_memcpy_d2h(void* dst, const void* src, int64_t bytes, void* stream);
int32_t sdcb200_stream_sync(void* stream);

/* ------------------------------------------------------------------ rolling
 * Series.rolling(window, min_periods).<op>()
 * replaces gen_sdc_pandas_series_rolling_impl / _minmax_impl / _ddof_impl
 * (sdc/datatypes/hpat_pandas_series_rolling_functions.py:586-733) and their
 * instantiations (:736-754).  Output is always float64[n] (:596).
 *   op       SDCB200_ROLL_*
 *   dtype    SDCB200_F64 or SDCB200_I64 input column (ints are read as numbers)
 *   win,minp already validated (0 <= minp <= win), minp defaulted to win by caller
 *   ddof     used by VAR / STD only
 *   halo     optional: the `halo_len` rows (same dtype as `in`) that precede in[0] in the global series
 *            (row-sharded multi-GPU: the previous rank's last win-1 rows, the
 *            analogue of the per-chunk prelude at :609-617); NULL / 0 otherwise
 *   row_offset  global row index of in[0] (count's "rows seen < minp" rule, :249-250)
 * Any window up to the length of the series, as in the reference: win <= 4097 runs the tiled kernel (one read and
 * one write per row); longer windows are cut at 2048-row chunk borders -- suffix of the first chunk, whole chunks
 * from a segment tree of chunk totals, prefix of the row's own chunk (four reads and one write per row).
 */
#define SDCB200_ROLL_SUM   0
#define SDCB200_ROLL_MEAN  1
#define SDCB200_ROLL_VAR   2
#define SDCB200_ROLL_STD   3
#define SDCB200_ROLL_COUNT 4
#define SDCB200_ROLL_MIN   5
#define SDCB200_ROLL_MAX   6

#define SDCB200_I8  0
#define SDCB200_U8  1
#define SDCB200_I16 2
#define SDCB200_U16 3
#define SDCB200_I32 4
#define SDCB200_U32 5
#define SDCB200_I64 6
#define SDCB200_U64 7
#define SDCB200_F32 8
#define SDCB200_F64 9
/* groupby key column of dense int64 group ids in [0, expected_groups); negative ids skip the row
 * (what sdcb200_str_factorize produces: the table of the aggregate is direct-addressed, no hashing) */
#define SDCB200_GID 100

int32_t sdcb200_rolling(int32_t op, int32_t dtype, const void* in, double* out, int64_t n,
                        int64_t win, int64_t minp, int64_t ddof,
                        const void* halo, int64_t halo_len, int64_t row_offset,
                        void* stream);
/* host-buffer form: chunks the column through pinned staging, overlapping H2D,
 * the kernel and D2H on three streams.                                          */
int32_t sdcb200_host_rolling(int32_t op, int32_t dtype, const void* in, double* out, int64_t n,
                             int64_t win, int64_t minp, int64_t ddof);
/* Row-sharded series, one process per GPU of an NVLink node: the win-1 rows in front of this rank's shard come from
 * the previous rank THROUGH PEER MEMORY INSIDE THE ROLLING KERNEL instead of a send / recv in front of it (the
 * analogue of the per-chunk prelude, hpat_pandas_series_rolling_functions.py:609-617, across GPUs).  Every rank owns
 * a mailbox of sdcb200_halo_mailbox_bytes() bytes (sdcb200_ipc_alloc, zeroed once, mapped by both neighbours with
 * sdcb200_ipc_open).  The kernel's first CTA stores this rank's last win-1 rows and a flag (epoch << 32 | rows) into
 * next_mailbox; the tile that needs the previous rank's rows runs last and waits for the flag in own_mailbox, then
 * acknowledges into prev_mailbox (a slot is reused only after the ack of epoch - 2).  next_mailbox / prev_mailbox are
 * NULL on the last / first rank.  epoch >= 1 and increases by one per call, identically on every rank.  Shards
 * shorter than win-1 rows forward rows they received themselves (exact for any shard lengths).  win - 1 must fit a
 * mailbox slot (4096 rows); for longer windows exchange the halo rows and call sdcb200_rolling.                 */
int64_t sdcb200_halo_mailbox_bytes(void);
int32_t sdcb200_rolling_linked(int32_t op, int32_t dtype, const void* in, double* out, int64_t n, int64_t win,
                               int64_t minp, int64_t ddof, int64_t row_offset, void* own_mailbox, void* next_mailbox,
                               void* prev_mailbox, uint32_t epoch, void* stream);

/* Series.rolling(win, minp).skew() / .kurt() / .cov(other, ddof) / .corr(other) on float64 DEVICE columns
 * (sdc/datatypes/hpat_pandas_series_rolling_functions.py: skew :408-431, :548-556; kurt :303-331, :520-536;
 * cov :261-300, :509-517, :925-992; corr :207-236, :487-506, :824-879).  A row takes part only when it is
 * finite in both series; `y` NULL means other = self.  Series of different lengths: the output has
 * max(n_x, n_y) rows, only rows below min(n_x, n_y) enter a window.  win <= 4097.                       */
#define SDCB200_ROLL_SKEW 7
#define SDCB200_ROLL_KURT 8
#define SDCB200_ROLL_COV 9
#define SDCB200_ROLL_CORR 10
int32_t sdcb200_rolling_moments(int32_t op, const double* x, int64_t n_x, const double* y, int64_t n_y,
                                double* out, int64_t win, int64_t minp, int64_t ddof, void* stream);

/* ------------------------------------------------------------------ sort / argsort / gather
 * In-place ascending sort with NaN last: replaces parallel_sort_<t> / parallel_stable_sort_<t>
 * (sdc/native/sort.cpp:119-121, sdc/native/stable_sort.cpp:308-310; exported at
 * sdc/native/module.cpp:33-61).  `dtype` is one of the ten SDCB200_* codes.                      */
int32_t sdcb200_sort(int32_t dtype, void* keys, int64_t n, void* stream);
/* Stable argsort: index_out[i] = row of the i-th element in the requested order; NaN last for
 * ascending AND descending, -0.0 == +0.0, ties in input order.  Replaces
 * parallel_[stable_]argsort_u64<t> (sdc/native/sort.cpp:94-104, stable_sort.cpp:283-293;
 * called from sdc/functions/sort.py:289-310).                                                     */
int32_t sdcb200_argsort(int32_t dtype, const void* keys, int64_t n, int32_t ascending, int64_t* index_out, void* stream);
/* dst[i] = src[idx[i]] for rows of elem_size 1/2/4/8 bytes: numpy_like.take
 * (sdc/functions/numpy_like.py:1359-1388) and `self._data[sorted_index]`
 * (sdc/datatypes/hpat_pandas_series_functions.py:3954-3955).                                      */
int32_t sdcb200_gather(int32_t elem_size, const void* src, const int64_t* idx, int64_t n, void* dst, void* stream);
int32_t sdcb200_host_sort(int32_t dtype, void* keys, int64_t n);
int32_t sdcb200_host_argsort(int32_t dtype, const void* keys, int64_t n, int32_t ascending, int64_t* index_out);

/* ------------------------------------------------------------------ groupby
 * df.groupby(key).{sum,mean,min,max,count,var,std,prod,median}() / Series.groupby(arr).<agg>():
 * replaces sdc_pandas_dataframe_groupby (sdc/datatypes/hpat_pandas_dataframe_functions.py:2993-3106),
 * merge_groupby_dicts_inplace and the per-group take + reduce loop
 * (sdc/datatypes/hpat_pandas_groupby_functions.py:58-73, 194-283) with one hash aggregate.
 *   key_dtype        SDCB200_I64 or SDCB200_F64 (NaN keys are skipped, -0.0 == +0.0), or SDCB200_GID
 *   cols/col_dtypes  ncols (<= 32) value columns, each SDCB200_F64 or SDCB200_I64, device pointers
 *   agg_mask         OR of SDCB200_AGG_*: every requested aggregate is produced for every column
 *   sort             1: result ordered by key (stable argsort, groupby_functions.py:209-210);
 *                    0: first-appearance order (the merged dict's order, :58-73)
 *   expected_groups  hint (<= 0: unknown) used to size the hash table / pick the shared-memory path
 * Result values are float64 except: count -> int64; sum/min/max/prod of an int64 column -> int64.
 * The result lives in a handle; copy keys / values out (device or host destination) and free it.  */
#define SDCB200_AGG_SUM   1u
#define SDCB200_AGG_MEAN  2u
#define SDCB200_AGG_MIN   4u
#define SDCB200_AGG_MAX   8u
#define SDCB200_AGG_COUNT 16u
#define SDCB200_AGG_VAR   32u
#define SDCB200_AGG_STD   64u
#define SDCB200_AGG_PROD  128u   /* nanprod: NaN skipped, all-NaN group -> 1; int64 columns multiply with wrap-around */
#define SDCB200_AGG_MEDIAN 256u  /* nanmedian of the group; always float64 */
int32_t sdcb200_groupby(int32_t key_dtype, const void* keys, int64_t n, int32_t ncols, const void* const* cols,
                        const int32_t* col_dtypes, uint32_t agg_mask, int32_t sort, int64_t ddof,
                        int64_t expected_groups, void* stream, int64_t* handle_out, int64_t* ngroups_out);
/* Asynchronous form for pipelines of small aggregates: _begin enqueues the work on `stream` and returns, _end waits
 * for the stream and yields the group count; the accessors below work on the handle after _end.  Requests on the
 * one-kernel low-cardinality path (at most one value column, sum / mean / min / max / count, group hint <= 4096 or 0,
 * n >= 4096) do not block in _begin; anything else completes inside it.  The input columns must stay valid until _end
 * (an optimistic attempt that overflowed is rerun there).  Same argument meaning as sdcb200_groupby.               */
int32_t sdcb200_groupby_begin(int32_t key_dtype, const void* keys, int64_t n, int32_t ncols, const void* const* cols,
                              const int32_t* col_dtypes, uint32_t agg_mask, int32_t sort, int64_t ddof,
                              int64_t expected_groups, void* stream, int64_t* handle_out);
int32_t sdcb200_groupby_end(int64_t handle, int64_t* ngroups_out);
int32_t sdcb200_groupby_keys(int64_t handle, void* dst, int32_t dst_is_host, void* stream);
int32_t sdcb200_groupby_values(int64_t handle, int32_t col, uint32_t agg, void* dst, int32_t dst_is_host, void* stream);
/* zero-copy access: device pointers of the handle's result, valid until sdcb200_groupby_free.
 * keys[ngroups]; vals is a matrix of rows of `stride` elements, row (rank(agg) * ncols + col) holds
 * that aggregate of that column in its first ngroups entries (rank = position of the aggregate among
 * the requested ones in ascending bit order).                                                       */
int32_t sdcb200_groupby_result(int64_t handle, void** keys, void** vals, int64_t* stride);
int32_t sdcb200_groupby_free(int64_t handle);
/* host-buffer form of sdcb200_groupby (the NumPy / NRT buffers a jitted SDC function holds, as every native of the
 * reference receives them): columns are copied in on two streams, the aggregate runs, the inputs are released; read
 * the result with sdcb200_groupby_keys / sdcb200_groupby_values(dst_is_host = 1), then sdcb200_groupby_free.      */
int32_t sdcb200_host_groupby(int32_t key_dtype, const void* keys, int64_t n, int32_t ncols, const void* const* cols,
                             const int32_t* col_dtypes, uint32_t agg_mask, int32_t sort, int64_t ddof,
                             int64_t expected_groups, int64_t* handle_out, int64_t* ngroups_out);

/* Composite ids for multi-key groupby / merge (the reference takes ONE `by` column,
 * sdc/datatypes/hpat_pandas_dataframe_functions.py:3066-3067): out[i] = a[i] * nb + b[i], `invalid` where either
 * id is negative.  The per-column ids are positions among the sorted distinct keys, so the composite id orders
 * the key tuples lexicographically.  Device pointers.                                                      */
int32_t sdcb200_combine_ids(const int64_t* a, const int64_t* b, int64_t nb, int64_t invalid, int64_t* out, int64_t n,
                            void* stream);

/* ------------------------------------------------------------------ join
 * pd.merge(left, right, on=key, how='inner' | 'left' | 'right' | 'outer') on one int64 / float64 key column.  The
 * reference snapshot has no merge overload (sdc/hiframes/join.py:34-43 is a stub; every test of
 * sdc/tests/test_join.py:51-394 is skipped): behaviour follows pandas.merge, the artefact follows
 * the reference's live indexer convention (_sdc_internal_join,
 * sdc/datatypes/common_functions.py:225-455): two int64 row-position arrays, -1 = no partner
 * (how='left' / 'right' / 'outer'), duplicate keys give the cross product.  inner / left: pairs come out
 * in left-row order; right: in right-row order; outer: the left join's pairs, then the unmatched right rows
 * in right-row order; the order of several matches of one row is unspecified.  Float keys: NaN == NaN,
 * -0.0 == +0.0 (pandas).  Build side (right; left for how='right') < 2^32 - 1 rows.  The result lives in a
 * handle.
 * how = SDCB200_JOIN_INNER | SDCB200_JOIN_ANY_ORDER: the caller takes the pairs as a row SET (what SURVEY 8(d) C3
 * compares, and all an aggregate or a further join downstream needs) -- they may come out in any order, and the order
 * may differ between two calls.  Keys with no dense range then take the partitioned join (both sides bucketed by hash
 * so that the table slice in use stays in L2; probe and emit in one pass, a tile taking its output range with one
 * atomic -- pairs come out roughly bucket by bucket, in no fixed order inside).  For dense keys the same single-pass
 * probe measured slower than the two passes that keep left-row order (1.12 against 1.04 ms on C3 -- the probe is bound
 * by L2 sector requests, not by its scan), so dense keys keep the two passes under this flag as well.          */
#define SDCB200_JOIN_INNER 0
#define SDCB200_JOIN_LEFT  1
#define SDCB200_JOIN_RIGHT 2   /* every right row: a left join with the sides swapped */
#define SDCB200_JOIN_ANY_ORDER 0x100   /* flag, how = INNER only */
#define SDCB200_JOIN_OUTER 3   /* the left join's pairs, then the unmatched right rows as (-1, row): the row set of
                                  _sdc_internal_join (sdc/datatypes/common_functions.py:225-455) */
int32_t sdcb200_join(int32_t how, int32_t key_dtype, const void* left_keys, int64_t nl, const void* right_keys, int64_t nr,
                     void* stream, int64_t* handle_out, int64_t* n_out);
/* The same join with the caller's ids in place of local row numbers: the result holds left_ids[l] / right_ids[r]
 * (device int64 columns, either may be NULL; -1 stays -1).  After a shuffle the rows carry the GLOBAL row numbers
 * they had before it (sdc_b200/distributed.py join_distributed); the mapping is folded into the probe's last kernel
 * (unique build keys) instead of two gather passes over the result.                                             */
int32_t sdcb200_join_ids(int32_t how, int32_t key_dtype, const void* left_keys, int64_t nl, const void* right_keys, int64_t nr,
                         const int64_t* left_ids, const int64_t* right_ids, void* stream, int64_t* handle_out, int64_t* n_out);
int32_t sdcb200_join_indexers(int64_t handle, int64_t* left_idx, int64_t* right_idx, int32_t dst_is_host, void* stream);
/* zero-copy access: device pointers of the handle's indexer arrays (n_out int64 each), valid until
 * sdcb200_join_free                                                                                  */
int32_t sdcb200_join_result(int64_t handle, int64_t** left_idx, int64_t** right_idx);
int32_t sdcb200_join_free(int64_t handle, void* stream);
/* host-buffer form of sdcb200_join; read the indexers with sdcb200_join_indexers(dst_is_host = 1). */
int32_t sdcb200_host_join(int32_t how, int32_t key_dtype, const void* left_keys, int64_t nl, const void* right_keys, int64_t nr,
                          int64_t* handle_out, int64_t* n_out);
/* dst[i] = idx[i] < 0 ? NaN : (double)src[idx[i]] -- payload gather of a left join
 * (setitem_arr_nan, sdc/hiframes/join.py:34-43); src_dtype SDCB200_I64 or SDCB200_F64.            */
int32_t sdcb200_take_nullable_f64(int32_t src_dtype, const void* src, const int64_t* idx, int64_t n, double* dst,
                                  void* stream);
/* pd.merge_asof(left, right, on=key) with the defaults direction='backward', allow_exact_matches=True (the shape of
 * sdc/tests/test_join.py:237-262, skipped there -- SURVEY 8f item 3): right_idx[i] = the LAST right row whose key is
 * <= left_keys[i], -1 when there is none.  Both key columns sorted ascending; device pointers.             */
int32_t sdcb200_asof_backward(int32_t key_dtype, const void* left_keys, int64_t nl, const void* right_keys, int64_t nr,
                              int64_t* right_idx, void* stream);

/* ------------------------------------------------------------------ string columns (str_arr)
 * Layout = the reference's StringArray (sdc/str_arr_type.py:33-101, sdc/native/str_ext/sdc_str_arr.hpp:160-176):
 * uint32 offsets[n + 1], uint8 chars[offsets[n]] (UTF-8, no terminators), validity bitmap LSB-first with
 * 1 = valid (NULL bitmap pointer = all valid).  All pointers below are DEVICE pointers.
 *
 * sdcb200_str_argsort: replaces `stable_argsort` (sdc_str_arr.hpp:565-593; bound at sdc/str_arr_ext.py:94-99):
 * unsigned byte-lexicographic order, a proper prefix first, ties in input order for both directions;
 * writes uint64 row positions like the reference.                                                   */
int32_t sdcb200_str_argsort(const uint8_t* chars, const uint32_t* offsets, int64_t n, int32_t ascending, uint64_t* index_out,
                            void* stream);
/* String keys -> exact dense group ids in [0, *ngroups_out), numbered in order of insertion (gid_out[i] = id
 * of row i's string, -1 for NA rows, which groupby skips: sdc/datatypes/hpat_pandas_dataframe_functions.py:3088-3089).
 * The ids feed sdcb200_groupby as a SDCB200_GID key column with expected_groups = *ngroups_out;
 * sdcb200_str_group_rows maps result ids back to one representative input row each (to materialise the key
 * strings with sdcb200_str_gather_*; table_cap = *table_cap_out); the table is released with
 * sdcb200_str_table_free.                                                                           */
int32_t sdcb200_str_factorize(const uint8_t* chars, const uint32_t* offsets, const uint8_t* bitmap, int64_t n,
                              int64_t expected_groups, int64_t* gid_out, int64_t* ngroups_out, int64_t* table_handle_out,
                              int64_t* table_cap_out, void* stream);
int32_t sdcb200_str_group_rows(int64_t table_handle, int64_t table_cap, const int64_t* gids, int64_t ng, int64_t* rows_out,
                               void* stream);
int32_t sdcb200_str_table_free(int64_t table_handle, void* stream);
/* take() on a str_arr (sdc/functions/numpy_like.py:1359-1388 for strings; convert_len_arr_to_offset,
 * sdc_str_arr.hpp:242-252): step 1 writes out_offsets[m + 1] and returns the char total, step 2 copies
 * the bytes and builds the bitmap.  rows[i] < 0 gives an empty NA string.                           */
int32_t sdcb200_str_gather_sizes(const uint32_t* offsets, const int64_t* rows, int64_t m, uint32_t* out_offsets,
                                 int64_t* total_chars, void* stream);
int32_t sdcb200_str_gather_chars(const uint8_t* chars, const uint32_t* offsets, const uint8_t* bitmap, const int64_t* rows,
                                 int64_t m, const uint32_t* out_offsets, uint8_t* out_chars, uint8_t* out_bitmap, void* stream);
/* String columns in the multi-GPU shuffle: the reference's wire format -- a LENGTH array plus the characters, offsets
 * rebuilt on the receiving side by a scan (sdc/shuffle_utils.py:277-297, convert_len_arr_to_offset
 * sdc/native/str_ext/sdc_str_arr.hpp:242-252).  The validity bit travels in bit 31 of the length.
 *   str_hash         out[i] = 64-bit hash of row i's bytes (NA rows: one fixed value): the routing key of a string column
 *   str_pack_lens    lens_out[i] = byte length | (NA ? 1 << 31 : 0)
 *   str_unpack_lens  lens -> offsets_out[n + 1], bitmap_out[(n + 7) / 8] (nullable), *total_chars (host)           */
int32_t sdcb200_str_hash(const uint8_t* chars, const uint32_t* offsets, const uint8_t* bitmap, int64_t n, uint64_t* out, void* stream);
int32_t sdcb200_str_pack_lens(const uint32_t* offsets, const uint8_t* bitmap, int64_t n, uint32_t* lens_out, void* stream);
int32_t sdcb200_str_unpack_lens(const uint32_t* lens, int64_t n, uint32_t* offsets_out, uint8_t* bitmap_out, int64_t* total_chars,
                                void* stream);
/* Short string keys as 64-bit words: a key column whose strings all fit one machine word goes through the numeric hash
 * groupby (sdcb200_groupby on int64 keys) without being factorized first.  The reference hashes the unboxed Python
 * strings into a numba Dict (hpat_pandas_dataframe_functions.py:3076-3097); the result is the same group set.
 *   str_profile       out3 (host) = {shortest length, longest length, number of NA rows} of the column
 *   str_pack_words    words_out[i] = bytes (little-endian, zero padded) | length << 56; every length <= 7
 *                     (a column of exactly-8-byte strings needs no packing: its chars buffer is the word column)
 *   str_unpack_words  result words -> offsets_out[m + 1] + chars_out (room for 8 m bytes); fixed8 != 0: every word is
 *                     an 8-byte string, else the length sits in the top byte                                          */
int32_t sdcb200_str_profile(const uint32_t* offsets, const uint8_t* bitmap, int64_t n, int64_t* out3, void* stream);
int32_t sdcb200_str_pack_words(const uint8_t* chars, const uint32_t* offsets, int64_t n, uint64_t* words_out, void* stream);
int32_t sdcb200_str_unpack_words(const uint64_t* words, int64_t m, int32_t fixed8, uint32_t* offsets_out, uint8_t* chars_out,
                                 int64_t* total_chars, void* stream);
/*   str_words_order   order_out[j] = the row of `words` that comes j-th in the reference's string order (unsigned bytes,
 *                     a proper prefix first: std::string `<`, sdc_str_arr.hpp:565-593) -- orders a groupby's result keys  */
int32_t sdcb200_str_words_order(const uint64_t* words, int64_t m, int32_t fixed8, int64_t* order_out, void* stream);
/* the reference's own symbol (HOST buffers, void return), exactly as sdc/str_arr_ext.py:94-99 declares it */
void stable_argsort(const char* data, const uint32_t* offsets, uint64_t len, int8_t ascending, uint64_t* result);

/* ------------------------------------------------------------------ CSV text -> numeric columns on the device
 * Replaces, for numeric tables, the reference's CSV ingest: pd.read_csv lowered to pyarrow.csv.read_csv on the host
 * (sdc/io/csv_ext.py:93-110, option mapping :174-272; sdc/native/arrow_reader.cpp), then boxing into NumPy.  Here the
 * file's bytes sit in device memory and are decoded there (csrc/csv.cu); number conversion is bit-identical to Arrow's
 * (int64 with overflow check; float64 correctly rounded, csrc/num_parse.cuh), empty lines are skipped, CRLF accepted,
 * Arrow's null spellings give NaN in a float64 column.
 *   csv_open      index the data rows of `data` (the bytes AFTER any header / skipped lines): *nrows_out rows
 *   csv_classify  kinds_out[c] (host) = OR over the rows of 1 int64 | 2 float64 | 4 null | 8 other -- Arrow's inference:
 *                 int64 if only 1 (and no 4), float64 if only 1 | 2 | 4, otherwise not a numeric column
 *   csv_parse     col_types[c]: 0 skip, 1 int64, 2 float64; out_cols[c] = device column of nrows elements.  A field
 *                 that does not convert makes the call fail with its row / column / reason in sdcb200_last_error()
 *   csv_close     releases the index (runs on the stream given to csv_open)                                         */
int32_t sdcb200_csv_open(const uint8_t* data, int64_t nbytes, int64_t* handle_out, int64_t* nrows_out, void* stream);
int32_t sdcb200_csv_classify(int64_t handle, int32_t ncols, int32_t delimiter, uint32_t* kinds_out);
int32_t sdcb200_csv_parse(int64_t handle, int32_t ncols, int32_t delimiter, const int32_t* col_types, void* const* out_cols);
int32_t sdcb200_csv_close(int64_t handle);

/* ------------------------------------------------------------------ multi-GPU hash shuffle
 * Partition step of the reference's (dead) MPI shuffle (sdc/shuffle_utils.py:119-163: per-row destination
 * and per-destination counts); the exchange -- c_alltoall of the counts and c_alltoallv per column,
 * sdc/distributed_api.py:439-466, sdc/transport/hpat_transport_single_process.cpp:106-119 -- is an NCCL
 * all-to-all issued by sdc_b200/distributed.py.  perm_out = row numbers grouped by destination rank
 * (rows of one destination keep their input order), counts_out[nparts] = rows per destination (int64).
 * Columns are packed for sending with sdcb200_gather(perm).                                           */
/* route: how a row finds its owner rank.
 *   SDCB200_ROUTE_HASH  destination = top bits of mix64(canonical key) scaled to nparts; the key travels unchanged.
 *   SDCB200_ROUTE_MOD   int64 keys only: destination = key mod nparts (Euclidean) and the key that travels is
 *                       floor(key / nparts): a dense key range stays dense on every rank, so the local join /
 *                       groupby keep their direct-addressed tables after the shuffle.  key = shipped * nparts + rank.
 * shipped_keys_out (nullable, n rows, input order): the key column as it travels.                              */
#define SDCB200_ROUTE_HASH 0
#define SDCB200_ROUTE_MOD 1
int32_t sdcb200_hash_partition(int32_t key_dtype, int32_t route, const void* keys, int64_t n, int32_t nparts,
                               int64_t* perm_out, int64_t* counts_out, void* shipped_keys_out, void* stream);
/* out[i] = cnt[i] ? sum[i] / cnt[i] : NaN: final step of a distributed mean (partials are exchanged).  */
int32_t sdcb200_combine_mean(const double* sum, const int64_t* cnt, int64_t n, double* out, void* stream);
/* Partial tables of a low-cardinality distributed groupby (the reference's per-chunk-dict-then-merge,
 * hpat_pandas_dataframe_functions.py:3080-3102, with GPUs as the chunks).  pack: `nrows` device rows of ng 8-byte
 * values (row 0 = the group keys) -> one int64 matrix dst[nrows][cap + 1], zero padded, dst[0][cap] = ng (an all-zero
 * body when ng > cap) -- what one all-gather moves.  combine: the gathered [nranks][nrows][cap + 1] block reduced over
 * the ranks into out[nrows][cap + 1], row r by ops[r] (SDCB200_PART_*: keys are copied from rank 0, float minima /
 * maxima skip NaN partials); flag[0] = 1 when some rank's key row or group count differs from rank 0's -- the caller
 * then merges by key instead -- flag[1] = rank 0's group count.  Table-sized work, one launch each.              */
#define SDCB200_PART_KEY 0
#define SDCB200_PART_SUM_F64 1
#define SDCB200_PART_SUM_I64 2
#define SDCB200_PART_MIN_F64 3
#define SDCB200_PART_MAX_F64 4
#define SDCB200_PART_MIN_I64 5
#define SDCB200_PART_MAX_I64 6
int32_t sdcb200_pack_partials(const void* const* rows, int32_t nrows, int64_t ng, int64_t cap, int64_t* dst, void* stream);
int32_t sdcb200_combine_partials(const int64_t* gathered, int32_t nranks, int32_t nrows, const int32_t* ops, int64_t cap,
                                 int64_t* out, int64_t* flag, void* stream);

/* ---- peer-memory shuffle over NVLink (one node, one process per GPU) --------------------------------------
 * Replaces the per-column alltoallv of the reference's shuffle (sdc/shuffle_utils.py:268-270,
 * sdc/distributed_api.py:439-466) with ONE pack-and-send kernel that stores rows straight into the
 * destination ranks' receive arenas (CUDA IPC + NVLink P2P stores).
 *   ipc_alloc / ipc_free      arena memory (cudaMalloc: exportable, unlike the stream-ordered pool)
 *   ipc_get_handle / ipc_open 64-byte CUDA IPC handle of an arena / a peer's arena mapped into this process
 * Column c of the rows this rank sends to rank d lands at peer_bases[d] + (c * arena_rows + dst_off[d] + k) * 8;
 * dst_off and peer_bases are HOST arrays, cols device pointers, rows 8 bytes.  The caller orders the send kernel
 * between two collectives on the same stream (counts exchange before, barrier after); see
 * sdc_b200/distributed.py: PeerShuffle.                                                                      */
int32_t sdcb200_ipc_alloc(void** p, int64_t bytes);
int32_t sdcb200_ipc_free(void* p);
int32_t sdcb200_ipc_get_handle(void* p, void* handle64);
int32_t sdcb200_ipc_open(const void* handle64, void** p);
int32_t sdcb200_ipc_close(void* p);
/* shuffle_plan hashes the keys once and leaves the per-tile
 * (4096 rows) exclusive offsets per destination in tile_off[ceil(n / 4096) * 8] (device, uint32) and the totals in
 * counts_out[nparts] (device); after the totals of every rank are known, shuffle_send re-hashes, ranks the rows of
 * each destination inside the tile in row order, stages them in shared memory and stores each destination's run
 * into that rank's arena.  cols[0] must be the key column.  Row order in the arenas = the NCCL path's.            */
int32_t sdcb200_shuffle_plan(int32_t key_dtype, int32_t route, const void* keys, int64_t n, int32_t nparts,
                             uint32_t* tile_off, int64_t* counts_out, void* stream);
/* iota_col: -1, or the index (>= 1) of a payload column that is generated instead of read -- its value for input
 * row r is iota_base + r (the global row id a join ships with every key; cols[iota_col] is ignored).        */
int32_t sdcb200_shuffle_send(int32_t key_dtype, int32_t route, int32_t ncols, const void* const* cols, int32_t iota_col,
                             int64_t iota_base, int64_t n, int32_t nparts, const uint32_t* tile_off,
                             const int64_t* dst_off, void* const* peer_bases, int64_t arena_rows, void* stream);
/* Range routing (sample sort: Series.sort_values over row-sharded columns): destination = the number of
 * `splitters` (host array of nparts - 1 ascending ORDERED IMAGES: int64 key ^ 2^63; float64 bits with the sign bit set
 * for non-negative values and all bits flipped for negative ones, -0.0 as +0.0) that are <= the key's image, reversed
 * when `descending`; NaN keys all go to rank nan_dest.  Keys travel unchanged; everything else as in plan / send.  */
int32_t sdcb200_shuffle_plan_range(int32_t key_dtype, const uint64_t* splitters, int32_t descending, int32_t nan_dest,
                                   const void* keys, int64_t n, int32_t nparts, uint32_t* tile_off, int64_t* counts_out,
                                   void* stream);
int32_t sdcb200_shuffle_send_range(int32_t key_dtype, const uint64_t* splitters, int32_t descending, int32_t nan_dest,
                                   int32_t ncols, const void* const* cols, int32_t iota_col, int64_t iota_base, int64_t n,
                                   int32_t nparts, const uint32_t* tile_off, const int64_t* dst_off, void* const* peer_bases,
                                   int64_t arena_rows, void* stream);
/* row-wise helpers of the distributed pass (device pointers):
 *   affine_i64   x[i] = x[i] * mul + add, in place; keep_negative != 0 leaves rows < 0 alone (the -1 "no partner"
 *                marker).  Shipped keys back to keys (mul = nparts, add = rank), local rows to global rows (add = offset).
 *   gather_ids   dst[i] = idx[i] < 0 ? -1 : src[idx[i]]: the local row numbers a join returns -> the global row ids
 *                that travelled with the rows.                                                                   */
int32_t sdcb200_affine_i64(int64_t* x, int64_t n, int64_t mul, int64_t add, int32_t keep_negative, void* stream);
int32_t sdcb200_gather_ids(const int64_t* src, const int64_t* idx, int64_t n, int64_t* dst, void* stream);

/* The reference's own entry points (sdc/native/module.cpp:31-92), same names and signatures,
 * HOST buffers, void return -- bind them exactly as sdc/functions/sort.py:39-72 does.  The
 * generic-comparator forms (parallel_sort(begin,len,size,compare) etc.) take a host callback and
 * cannot run on a GPU: not provided.                                                              */
#define SDCB200_DECL_REF_SORT(T)                                                         \
    void parallel_sort_##T(void* begin, uint64_t len);                                   \
    void parallel_stable_sort_##T(void* begin, uint64_t len);                            \
    void parallel_argsort_u64##T(void* index, void* begin, uint64_t len, uint8_t ascending);        \
    void parallel_stable_argsort_u64##T(void* index, void* begin, uint64_t len, uint8_t ascending);
SDCB200_DECL_REF_SORT(i8)
SDCB200_DECL_REF_SORT(u8)
SDCB200_DECL_REF_SORT(i16)
SDCB200_DECL_REF_SORT(u16)
SDCB200_DECL_REF_SORT(i32)
SDCB200_DECL_REF_SORT(u32)
SDCB200_DECL_REF_SORT(i64)
SDCB200_DECL_REF_SORT(u64)
SDCB200_DECL_REF_SORT(f32)
SDCB200_DECL_REF_SORT(f64)
void set_number_of_threads(uint64_t threads);

#ifdef __cplusplus
}
#endif
#endif /* SDC_B200_H */
