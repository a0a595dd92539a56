"""sort / argsort / Series.sort_values on the B200 backend.

Host-side mirror of
  * numpy_like.sort / argsort dispatch ......... sdc/functions/numpy_like.py:1197-1278
  * parallel_[stable_]sort / argsort bindings ... sdc/functions/sort.py:243-340
  * Series.sort_values ........................... sdc/datatypes/hpat_pandas_series_functions.py:3853-3959
The ordering runs in libsdc_b200.so (csrc/sort.cu: onesweep radix sort).  Both `kind`s map to the
same stable kernel (a stable order is a valid 'quicksort' order).  torch is used only to hold
device buffers and move bytes.
"""
import numpy as np

from . import native

_KINDS = (None, "quicksort", "mergesort")


def _torch():
    import torch
    return torch


def _code(dtype_name):
    try:
        return native.DTYPE_CODES[dtype_name]
    except KeyError:
        raise TypeError("sort: unsupported dtype %s" % dtype_name)


def _np_name(t):
    return str(t.dtype).replace("torch.", "")


def sort_device(col):
    """In-place ascending sort (NaN last) of a contiguous 1-D CUDA tensor."""
    torch = _torch()
    lib = native.load()
    assert col.is_cuda and col.dim() == 1 and col.is_contiguous()
    with torch.cuda.device(col.device):
        native.check(lib.sdcb200_sort(_code(_np_name(col)), col.data_ptr(), col.shape[0],
                                      native.current_stream_ptr()))
    return col


def argsort_device(col, ascending=True):
    """Stable argsort (NaN last in both directions) of a 1-D CUDA tensor -> int64 CUDA tensor."""
    torch = _torch()
    lib = native.load()
    assert col.is_cuda and col.dim() == 1 and col.is_contiguous()
    out = torch.empty(col.shape[0], dtype=torch.int64, device=col.device)
    with torch.cuda.device(col.device):
        native.check(lib.sdcb200_argsort(_code(_np_name(col)), col.data_ptr(), col.shape[0],
                                         1 if ascending else 0, out.data_ptr(), native.current_stream_ptr()))
    return out


def gather_device(col, idx, out=None):
    """out[i] = col[idx[i]] (numpy_like.take) for 1/2/4/8-byte rows."""
    torch = _torch()
    lib = native.load()
    assert col.is_cuda and idx.is_cuda and idx.dtype == torch.int64 and col.is_contiguous() and idx.is_contiguous()
    if out is None:
        out = torch.empty(idx.shape[0], dtype=col.dtype, device=col.device)
    with torch.cuda.device(col.device):
        native.check(lib.sdcb200_gather(col.element_size(), col.data_ptr(), idx.data_ptr(), idx.shape[0],
                                        out.data_ptr(), native.current_stream_ptr()))
    return out


def _check_kind(kind, who="sort"):
    if kind not in _KINDS:
        raise ValueError("Unsupported value of 'kind' parameter")


def sort(a, kind=None):
    """numpy_like.sort: sorts the NumPy array IN PLACE (ascending, NaN last) and returns it."""
    _check_kind(kind)
    lib = native.load()
    a_c = a if a.flags.c_contiguous else np.ascontiguousarray(a)
    native.check(lib.sdcb200_host_sort(_code(a_c.dtype.name), a_c.ctypes.data, a_c.shape[0]))
    if a_c is not a:
        a[...] = a_c
    return a


def argsort(a, kind=None, ascending=True):
    """numpy_like.argsort: int64 positions; stable, NaN last in both directions."""
    _check_kind(kind)
    lib = native.load()
    a_c = np.ascontiguousarray(a)
    out = np.empty(a_c.shape[0], dtype=np.int64)
    native.check(lib.sdcb200_host_argsort(_code(a_c.dtype.name), a_c.ctypes.data, a_c.shape[0],
                                          1 if ascending else 0, out.ctypes.data))
    return out


def sort_values_indexer_device(col, ascending=True, na_position="last"):
    """int64 positions of Series.sort_values: non-NaN rows in order, NaN block first or last."""
    torch = _torch()
    idx = argsort_device(col, ascending)
    if na_position == "first" and col.dtype.is_floating_point:
        n_nan = int(torch.isnan(col).sum().item())      # plumbing: a count, not part of the ordering
        if n_nan:
            idx = torch.roll(idx, n_nan)
    return idx


def series_sort_values(series, axis=0, ascending=True, inplace=False, kind="quicksort", na_position="last"):
    """`Series.sort_values` for a numeric pandas Series (reference :3853-3959)."""
    import pandas as pd
    torch = _torch()
    if axis not in (0, "index"):
        raise ValueError("Method sort_values(). The object axis\n expected: 0")
    if inplace is not False:
        raise TypeError("Method sort_values(). Unsupported parameters. Given inplace: {}".format(inplace))
    if kind not in _KINDS:
        raise ValueError("Method sort_values(). Unsupported parameter. Given kind != 'quicksort', 'mergesort'")
    if na_position not in ("last", "first"):
        raise ValueError("Method sort_values(). Unsupported parameter. Given na_position != 'last', 'first'")
    values = series.to_numpy()
    if values.dtype.kind not in "iuf":
        from . import str_arr
        return str_arr.series_sort_values_str(series, ascending, na_position)
    col = torch.from_numpy(np.ascontiguousarray(values)).cuda()
    idx = sort_values_indexer_device(col, bool(ascending), na_position)
    data = gather_device(col, idx).cpu().numpy()
    index_values = series.index.to_numpy()
    if index_values.dtype.kind in "iuf" and index_values.dtype.itemsize in (1, 2, 4, 8):
        icol = torch.from_numpy(np.ascontiguousarray(index_values)).cuda()
        new_index = pd.Index(gather_device(icol, idx).cpu().numpy(), name=series.index.name)
    else:
        new_index = series.index.take(idx.cpu().numpy())
    return pd.Series(data, index=new_index, name=series.name)
