"""pd.read_csv for numeric tables, decoded on the device.

Host-side mirror of the reference's CSV ingest -- `pd.read_csv(filepath_or_buffer, sep, delimiter, names, usecols,
dtype, skiprows)` as sdc lowers it (sdc/io/csv_ext.py:93-110 do_read_csv, option mapping :174-272; the overload in
sdc/datatypes/hpat_pandas_functions.py) -- with the same argument meaning:
  * names=None: the first line (after `skiprows` lines) holds the column names; names given: every line is data
    (csv_ext.py:174-189);
  * usecols: column names or positions to keep (csv_ext.py:199-216);
  * dtype: one dtype or {column: dtype}; columns without one are inferred the way Arrow infers them -- int64 if every
    field is an integer, float64 if the rest are numbers or missing (csv_ext.py:217-231).
The bytes of the file are copied to the GPU once (pinned staging buffer) and everything row-proportional runs in
libsdc_b200.so (csrc/csv.cu): row index, type inference, number conversion (bit-identical to Arrow's).  A column that
is not numeric (strings, quoted fields, dates) makes `read_csv_device` raise; `DeviceFrame.read_csv(engine="auto")`
hands such files to Arrow's reader, whose string buffers are the str_arr layout.
"""
import ctypes as ct
import os

import numpy as np

from . import native

_KIND_INT, _KIND_FLOAT, _KIND_NULL, _KIND_OTHER = 1, 2, 4, 8
_TYPE_CODE = {"int64": 1, "float64": 2}


class NotNumericCSV(ValueError):
    """The file has a column the device reader does not decode (strings, quoted fields)."""


def _torch():
    import torch
    return torch


_CHUNK = 32 << 20


def _stage_to_device(filepath_or_buffer, dev):
    """File (or bytes) -> (uint8 CUDA tensor, the first MiB as host bytes, byte count).  The bytes pass through a pinned
    staging buffer in 32 MiB pieces: a few threads fill the pieces (file reads / memcpy release the GIL) while the pieces
    already filled are on their way to the device."""
    import concurrent.futures as cf
    torch = _torch()
    fd = None
    if isinstance(filepath_or_buffer, (bytes, bytearray, memoryview)):
        raw = np.frombuffer(filepath_or_buffer, dtype=np.uint8)
        n = raw.shape[0]
    elif hasattr(filepath_or_buffer, "read"):
        data = filepath_or_buffer.read()
        return _stage_to_device(data.encode("utf-8") if isinstance(data, str) else data, dev)
    else:
        fd = os.open(filepath_or_buffer, os.O_RDONLY)
        n = os.fstat(fd).st_size
    try:
        pinned = torch.empty(max(1, n), dtype=torch.uint8, pin_memory=True)
        out = torch.empty(max(16, n), dtype=torch.uint8, device=dev)
        host = pinned.numpy()

        def fill(a, b):
            if fd is None:
                np.copyto(host[a:b], raw[a:b])
            else:
                view, got = memoryview(host)[a:b], 0
                while got < b - a:
                    k = os.preadv(fd, [view[got:]], a + got)
                    if k <= 0:
                        raise IOError("read_csv: short read")
                    got += k
            return a, b
        spans = [(a, min(n, a + _CHUNK)) for a in range(0, n, _CHUNK)]
        with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(spans)))) as pool:
            for a, b in pool.map(lambda ab: fill(*ab), spans):            # in file order, as they complete
                out[a:b].copy_(pinned[a:b], non_blocking=True)
        head = bytes(host[:min(n, 1 << 20)])
        # the pinned buffer goes back to torch's caching host allocator only after the copies that read it are done
        torch.cuda.current_stream(dev).synchronize()
    finally:
        if fd is not None:
            os.close(fd)
    return out, head, n


def _header_split(host_bytes, nbytes, skiprows, want_header, delimiter):
    """Byte offset where the data rows begin and the header's names (host work on the first lines only)."""
    pos = 0
    for _ in range(int(skiprows or 0)):
        nl = host_bytes.find(b"\n", pos, nbytes)
        if nl < 0:
            return nbytes, None
        pos = nl + 1
    names = None
    if want_header:
        # Arrow skips empty lines in front of the header too
        while pos < nbytes:
            nl = host_bytes.find(b"\n", pos, nbytes)
            end = nbytes if nl < 0 else nl
            line = host_bytes[pos:end].rstrip(b"\r")
            pos = nbytes if nl < 0 else nl + 1
            if line:
                names = [x.decode("utf-8") for x in line.split(delimiter)]
                if any('"' in x for x in names):
                    names = [x[1:-1] if len(x) >= 2 and x[0] == '"' and x[-1] == '"' else x for x in names]
                break
    return pos, names


def read_csv_device(filepath_or_buffer, sep=",", delimiter=None, names=None, usecols=None, dtype=None, skiprows=None):
    """-> ({column name: CUDA tensor (int64 / float64)} in file order restricted to usecols, number of rows)."""
    torch = _torch()
    lib = native.load()
    if delimiter is None:
        delimiter = sep
    if not isinstance(delimiter, str) or len(delimiter.encode("utf-8")) != 1:
        raise ValueError("read_csv: the delimiter must be one byte")
    dbyte = delimiter.encode("utf-8")
    dev = torch.device("cuda", torch.cuda.current_device())
    data, head, nbytes = _stage_to_device(filepath_or_buffer, dev)
    start, file_names = _header_split(head, len(head), skiprows, names is None, dbyte)     # column names live in the first lines
    if names is None and file_names is None:
        return {}, 0
    col_names = list(names) if names is not None else file_names
    ncols = len(col_names)
    handle, nrows = ct.c_int64(0), ct.c_int64(0)
    stream = native.current_stream_ptr()
    native.check(lib.sdcb200_csv_open(data.data_ptr() + start, nbytes - start, ct.byref(handle), ct.byref(nrows), stream))
    try:
        n = nrows.value
        # which columns, and as what
        if usecols is None:
            keep = list(range(ncols))
        else:
            keep = sorted({(col_names.index(u) if isinstance(u, str) else int(u)) for u in usecols})
        want = {}
        if dtype is not None:
            items = dtype.items() if isinstance(dtype, dict) else [(col_names[i], dtype) for i in keep]
            for k, v in items:
                name = np.dtype(v).name
                if name not in _TYPE_CODE:
                    raise NotNumericCSV("read_csv: column %r: dtype %s is not decoded on the device" % (k, name))
                want[k] = name
        types = (ct.c_int32 * ncols)()
        if n > 0 and any(col_names[i] not in want for i in keep):
            kinds = (ct.c_uint32 * ncols)()
            native.check(lib.sdcb200_csv_classify(handle.value, ncols, dbyte[0], kinds))
        else:
            kinds = None
        outs = (ct.c_void_p * ncols)()
        cols = {}
        for i in keep:
            name = col_names[i]
            if name in want:
                kind = want[name]
            else:
                k = kinds[i] if kinds is not None else _KIND_NULL
                if k & _KIND_OTHER:
                    raise NotNumericCSV("read_csv: column %r is not numeric" % name)
                kind = "int64" if k == _KIND_INT else "float64"      # a column of missing values only: float64 NaN (pandas' rule)
            t = torch.empty(n, dtype=torch.int64 if kind == "int64" else torch.float64, device=dev)
            cols[name] = t
            types[i] = _TYPE_CODE[kind]
            outs[i] = t.data_ptr()
        if n > 0:
            native.check(lib.sdcb200_csv_parse(handle.value, ncols, dbyte[0], types, outs))
    finally:
        native.check(lib.sdcb200_csv_close(handle.value))
    return cols, n
