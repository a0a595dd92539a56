"""String columns on the B200 backend: the reference's str_arr layout held in device memory.

Host-side mirror of
  * StringArrayType / payload (offsets uint32[n+1], chars uint8[total], validity bitmap LSB-first,
    1 = valid; NA strings stored empty) ............ sdc/str_arr_type.py:33-101, sdc/native/str_ext/sdc_str_arr.hpp:160-176, 310-314, 389-393
  * str_arr_stable_argosort -> hstr_ext.stable_argsort  sdc/str_arr_ext.py:1468-1476, sdc_str_arr.hpp:565-593
  * Series.sort_values on strings (NaN split, argsort of the rest, na_position) sdc/datatypes/hpat_pandas_series_functions.py:3924-3957
  * groupby with string keys (None keys skipped) ...... sdc/datatypes/hpat_pandas_dataframe_functions.py:3076-3104
Ordering, hashing and gathering run in libsdc_b200.so (csrc/strings.cu, groupby.cu); the Python <-> str_arr
conversion below is boxing (the reference does it element by element under the GIL, sdc_str_arr.hpp:323-477).
"""
import ctypes as ct

import numpy as np

from . import native


def _torch():
    import torch
    return torch


class StringArray:
    """offsets / chars / bitmap as NumPy arrays (host) or torch CUDA tensors (device)."""

    def __init__(self, offsets, chars, bitmap, n):
        self.offsets, self.chars, self.bitmap, self.n = offsets, chars, bitmap, int(n)

    # ---- boxing ----
    @classmethod
    def from_pylist(cls, values):
        n = len(values)
        enc = [b"" if v is None or (isinstance(v, float) and v != v) else str(v).encode("utf-8") for v in values]
        lens = np.fromiter((len(e) for e in enc), dtype=np.int64, count=n)
        total = int(lens.sum())
        if total >= 2**32:
            raise ValueError("string array exceeds the uint32 offset range")          # sdc_str_arr.hpp:206
        offsets = np.zeros(n + 1, dtype=np.uint32)
        np.cumsum(lens, out=offsets[1:], dtype=np.uint64) if n else None
        chars = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if total else np.zeros(0, dtype=np.uint8)
        valid = np.fromiter((not (v is None or (isinstance(v, float) and v != v)) for v in values), dtype=bool, count=n)
        bitmap = np.packbits(valid, bitorder="little") if n else np.zeros(0, dtype=np.uint8)
        return cls(offsets, chars, bitmap, n)

    @classmethod
    def from_arrow(cls, arr):
        """pyarrow string / large_string array (or chunked array) -> str_arr WITHOUT touching Python strings: Arrow's
        layout is the reference's (offsets + UTF-8 data + validity bitmap, LSB-first, 1 = valid --
        sdc/str_arr_type.py:69-70), so the buffers are viewed, rebased to offset 0 and narrowed to uint32.  The
        reference's own CSV path hands Arrow string columns over the same way (sdc/native/arrow_reader.cpp)."""
        import pyarrow as pa
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
        if pa.types.is_dictionary(arr.type):
            arr = arr.dictionary_decode()
        if not (pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type)):
            raise TypeError("StringArray.from_arrow: expected a string array, got %s" % arr.type)
        n = len(arr)
        wide = pa.types.is_large_string(arr.type)
        vbuf, obuf, dbuf = arr.buffers()
        odt = np.int64 if wide else np.int32
        offs = np.frombuffer(obuf, dtype=odt, count=arr.offset + n + 1)[arr.offset:] if n else np.zeros(1, dtype=odt)
        lo, hi = int(offs[0]), int(offs[-1])
        if hi - lo >= 2**32:
            raise ValueError("string array exceeds the uint32 offset range")          # sdc_str_arr.hpp:206
        offsets = (offs - lo).astype(np.uint32)
        chars = np.frombuffer(dbuf, dtype=np.uint8, count=hi)[lo:hi].copy() if (dbuf is not None and hi > lo) else np.zeros(0, dtype=np.uint8)
        if vbuf is None or arr.null_count == 0:
            bitmap = np.packbits(np.ones(n, dtype=np.uint8), bitorder="little") if n else np.zeros(0, dtype=np.uint8)
        else:
            bits = np.unpackbits(np.frombuffer(vbuf, dtype=np.uint8), bitorder="little")[arr.offset:arr.offset + n]
            bitmap = np.packbits(bits, bitorder="little")
            if bits.size and not bits.all():
                # NA strings are stored empty (sdc_str_arr.hpp:389-393); Arrow may keep bytes under a null slot
                lens = np.diff(offsets.astype(np.int64))
                if (lens[bits == 0] != 0).any():
                    keep = np.repeat(bits.astype(bool), lens)
                    chars = chars[keep]
                    lens = np.where(bits.astype(bool), lens, 0)
                    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
        return cls(offsets, chars, bitmap, n)

    def to_arrow(self):
        """-> pyarrow.StringArray sharing the layout (one copy of each buffer to the host)."""
        import pyarrow as pa
        h = self.cpu()
        valid = h.valid_mask()
        vb = None if valid.all() else pa.py_buffer(np.packbits(valid, bitorder="little").tobytes())
        return pa.StringArray.from_buffers(h.n, pa.py_buffer(h.offsets.astype(np.int32).tobytes()), pa.py_buffer(h.chars.tobytes()),
                                           vb, -1 if vb is not None else 0)

    @classmethod
    def from_fixed_width(cls, codes_u8):
        """n x w uint8 matrix of characters -> str_arr without touching Python strings (benchmarks)."""
        n, w = codes_u8.shape
        offsets = (np.arange(n + 1, dtype=np.uint64) * w).astype(np.uint32)
        return cls(offsets, np.ascontiguousarray(codes_u8).reshape(-1), np.full((n + 7) // 8, 0xFF, dtype=np.uint8), n)

    def is_device(self):
        return not isinstance(self.offsets, np.ndarray)

    def cuda(self):
        torch = _torch()
        if self.is_device():
            return self
        pad = lambda a: a if a.size else np.zeros(1, dtype=a.dtype)      # keep data_ptr() valid for empty parts
        return StringArray(torch.from_numpy(self.offsets.view(np.int32)).cuda(), torch.from_numpy(pad(self.chars)).cuda(),
                           torch.from_numpy(pad(self.bitmap)).cuda(), self.n)

    def cpu(self):
        if not self.is_device():
            return self
        return StringArray(self.offsets.cpu().numpy().view(np.uint32), self.chars.cpu().numpy(), self.bitmap.cpu().numpy(), self.n)

    def valid_mask(self):
        h = self.cpu()
        return np.unpackbits(h.bitmap, bitorder="little")[:h.n].astype(bool) if h.n else np.zeros(0, dtype=bool)

    def to_pylist(self):
        h = self.cpu()
        valid = h.valid_mask()
        buf = h.chars.tobytes()
        o = h.offsets
        return [buf[o[i]:o[i + 1]].decode("utf-8") if valid[i] else None for i in range(h.n)]

    # ---- kernels ----
    def argsort(self, ascending=True):
        """Stable argsort in the reference's string order (NA rows sort as empty strings) -> CUDA int64."""
        torch = _torch()
        lib = native.load()
        d = self.cuda()
        out = torch.empty(d.n, dtype=torch.int64, device=d.offsets.device)
        with torch.cuda.device(d.offsets.device):
            native.check(lib.sdcb200_str_argsort(d.chars.data_ptr(), d.offsets.data_ptr(), d.n, 1 if ascending else 0,
                                                 out.data_ptr(), native.current_stream_ptr()))
        return out

    def take(self, rows):
        """rows: CUDA int64 tensor; rows < 0 give NA.  -> device StringArray."""
        torch = _torch()
        lib = native.load()
        d = self.cuda()
        m = rows.shape[0]
        dev = d.offsets.device
        out_off = torch.empty(m + 1, dtype=torch.int32, device=dev)
        total = ct.c_int64(0)
        with torch.cuda.device(dev):
            stream = native.current_stream_ptr()
            native.check(lib.sdcb200_str_gather_sizes(d.offsets.data_ptr(), rows.data_ptr(), m, out_off.data_ptr(),
                                                      ct.byref(total), stream))
            out_chars = torch.empty(max(1, total.value), dtype=torch.uint8, device=dev)
            out_bm = torch.empty(max(1, (m + 7) // 8), dtype=torch.uint8, device=dev)
            native.check(lib.sdcb200_str_gather_chars(d.chars.data_ptr(), d.offsets.data_ptr(), d.bitmap.data_ptr(),
                                                      rows.data_ptr(), m, out_off.data_ptr(), out_chars.data_ptr(),
                                                      out_bm.data_ptr(), stream))
        return StringArray(out_off, out_chars[:total.value] if total.value else out_chars[:0], out_bm[:(m + 7) // 8], m)

    def profile(self):
        """-> (shortest length, longest length, number of NA rows): one pass over offsets + bitmap, kept on the column
        (a str_arr is immutable: the reference never writes into one after it is built)."""
        prof = getattr(self, "_profile", None)
        if prof is None:
            torch = _torch()
            lib = native.load()
            d = self.cuda()
            out = (ct.c_int64 * 3)()
            with torch.cuda.device(d.offsets.device):
                native.check(lib.sdcb200_str_profile(d.offsets.data_ptr(), d.bitmap.data_ptr(), d.n, out,
                                                     native.current_stream_ptr()))
            prof = self._profile = (int(out[0]), int(out[1]), int(out[2]))
        return prof

    def key_words(self):
        """The column as int64 key words for the numeric hash groupby, or None when its strings do not all fit one:
        -> (int64 CUDA tensor[n], fixed8).  Exactly-8-byte strings: a zero-copy view of the chars; at most 7 bytes:
        bytes | length << 56 (csrc/strings.cu str_pack_words).  NA rows would have to be skipped: such columns are
        factorized instead."""
        torch = _torch()
        d = self.cuda()
        if d.n == 0 or d.offsets.data_ptr() % 16:
            return None
        mn, mx, n_na = d.profile()
        if n_na:
            return None
        if mn == 8 and mx == 8 and d.chars.data_ptr() % 16 == 0:
            return d.chars[:8 * d.n].view(torch.int64), True
        if mx <= 7 and d.chars.data_ptr() % 8 == 0:
            words = torch.empty(d.n, dtype=torch.int64, device=d.offsets.device)
            with torch.cuda.device(d.offsets.device):
                native.check(native.load().sdcb200_str_pack_words(d.chars.data_ptr(), d.offsets.data_ptr(), d.n,
                                                                  words.data_ptr(), native.current_stream_ptr()))
            return words, False
        return None

    @classmethod
    def from_key_words(cls, words, fixed8):
        """Result key words of the numeric groupby back into a device str_arr (no NA rows)."""
        torch = _torch()
        lib = native.load()
        m = words.shape[0]
        dev = words.device
        off = torch.empty(m + 1, dtype=torch.int32, device=dev)
        chars = torch.empty(max(1, 8 * m), dtype=torch.uint8, device=dev)
        total = ct.c_int64(0)
        with torch.cuda.device(dev):
            native.check(lib.sdcb200_str_unpack_words(words.data_ptr() if m else None, m, 1 if fixed8 else 0, off.data_ptr(),
                                                      chars.data_ptr(), ct.byref(total), native.current_stream_ptr()))
        return cls(off, chars[:total.value], torch.full((max(1, (m + 7) // 8),), 0xFF, dtype=torch.uint8, device=dev)[:(m + 7) // 8], m)

    def factorize(self, expected_groups=0):
        """-> (gid int64 CUDA tensor with -1 for NA rows, n_groups, table handle, table capacity)."""
        torch = _torch()
        lib = native.load()
        d = self.cuda()
        gid = torch.empty(d.n, dtype=torch.int64, device=d.offsets.device)
        ng, handle, cap = ct.c_int64(0), ct.c_int64(0), ct.c_int64(0)
        with torch.cuda.device(d.offsets.device):
            native.check(lib.sdcb200_str_factorize(d.chars.data_ptr(), d.offsets.data_ptr(), d.bitmap.data_ptr(), d.n,
                                                   int(expected_groups), gid.data_ptr(), ct.byref(ng), ct.byref(handle),
                                                   ct.byref(cap), native.current_stream_ptr()))
        return gid, ng.value, handle.value, cap.value


def groupby_str_device(keys, cols, aggs, sort=True, ddof=1, expected_groups=0, short_keys=True):
    """Hash groupby with a string key column: factorize -> numeric groupby on the ids -> key strings.

    keys: StringArray (device);  cols: list of CUDA tensors -> (result key StringArray, {agg: [tensor per column]})
    sort=True orders the groups by the reference's string order, sort=False by first appearance.
    short_keys=False keeps a column of one-word strings on the factorize path as well (A/B measurements, tests)."""
    torch = _torch()
    lib = native.load()
    from .groupby import groupby_device
    from .sort import gather_device
    d = keys.cuda()
    dev = d.offsets.device
    kw = d.key_words() if short_keys else None
    if kw is not None:
        # strings that fit one word: the numeric hash groupby on the words themselves (no factorize pass, no id column);
        # sort=True orders the few result rows by string order below, sort=False keeps first appearance
        words, fixed8 = kw
        gk, res = groupby_device(words, cols, aggs, bool(sort), ddof, expected_groups)
        if sort and gk.shape[0] > 1:
            # the aggregate ordered the groups by the words' integer value; string order is the order of their
            # big-endian image (the length below it), sorted on the few result rows
            order = torch.empty(gk.shape[0], dtype=torch.int64, device=dev)
            with torch.cuda.device(dev):
                native.check(lib.sdcb200_str_words_order(gk.data_ptr(), gk.shape[0], 1 if fixed8 else 0, order.data_ptr(),
                                                         native.current_stream_ptr()))
            gk = gather_device(gk, order)
            res = {a: [gather_device(t, order) for t in ts] for a, ts in res.items()}
        return StringArray.from_key_words(gk, fixed8), res
    gid, ng, handle, cap = d.factorize(expected_groups)
    try:
        # ids are dense, in [0, ng); NA rows carry -1 and are skipped by the aggregate (SDCB200_GID keys: a
        # direct-addressed table exactly as large as the result).
        # sort=True reorders by string order below, so the id order of the aggregate is as good as any (and cheaper
        # than tracking first rows); sort=False needs the first-appearance order from the aggregate itself
        gk, res = groupby_device(gid, cols, aggs, bool(sort), ddof, gid_range=max(ng, 1))
        rows = torch.empty(gk.shape[0], dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            native.check(lib.sdcb200_str_group_rows(handle, cap, gk.data_ptr(), gk.shape[0], rows.data_ptr(),
                                                    native.current_stream_ptr()))
    finally:
        with torch.cuda.device(dev):
            native.check(lib.sdcb200_str_table_free(handle, native.current_stream_ptr()))
    rkeys = d.take(rows)
    if sort and rkeys.n > 1:
        order = rkeys.argsort(True)
        rkeys = rkeys.take(order)
        res = {a: [gather_device(t, order) for t in ts] for a, ts in res.items()}
    return rkeys, res


def series_sort_values_str(series, ascending=True, na_position="last"):
    """Series.sort_values for a string Series (reference :3924-3957): None/NaN rows are split off, the rest
    is argsorted with the native stable string sort, the NA block goes last or first."""
    import pandas as pd
    torch = _torch()
    values = series.to_numpy(dtype=object)
    sa = StringArray.from_pylist(list(values))
    valid = sa.valid_mask()
    good = np.flatnonzero(valid)
    bad = np.flatnonzero(~valid)
    sub = StringArray.from_pylist([values[i] for i in good]).cuda()
    order = sub.argsort(bool(ascending)).cpu().numpy()
    pos = good[order]
    idx = np.concatenate([bad, pos]) if na_position == "first" else np.concatenate([pos, bad])
    return pd.Series(values[idx], index=series.index.take(idx), name=series.name, dtype=series.dtype)
