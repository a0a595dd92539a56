"""Series / DataFrame `.rolling(window, min_periods).<op>()` on the B200 backend.

Host-side mirror of the reference's rolling interface:
  * argument checks and messages ........ sdc/datatypes/hpat_pandas_rolling_types.py:108-176
  * Series methods ....................... sdc/datatypes/hpat_pandas_series_rolling_functions.py
        count :886-892, max :1022-1028, mean :1031-1037, min :1049-1055,
        std :1114-1123, sum :1126-1132, var :1135-1144
  * DataFrame = per-column Series rolling  sdc/datatypes/hpat_pandas_dataframe_rolling_functions.py:167-189
The arithmetic runs in libsdc_b200.so (csrc/rolling.cu); nothing here computes a window.
"""
import ctypes as ct

import numpy as np

from . import native

_SUPPORTED_IN = ("float64", "int64")


def check_rolling_args(window, min_periods=None, center=False, win_type=None, on=None, axis=0, closed=None):
    """Validate like the reference's `sdc_pandas_rolling_impl`; return the effective min_periods."""
    if not isinstance(window, (int, np.integer)) or isinstance(window, bool):
        raise TypeError('Method rolling(). The object window\n given: {}\n expected: int'.format(type(window).__name__))
    if min_periods is not None and (not isinstance(min_periods, (int, np.integer)) or isinstance(min_periods, bool)):
        raise TypeError('Method rolling(). The object min_periods\n given: {}\n expected: None, int'.format(
            type(min_periods).__name__))
    if window < 0:
        raise ValueError('window must be non-negative')
    minp = window if min_periods is None else min_periods
    if minp < 0:
        raise ValueError('min_periods must be >= 0')
    if minp > window:
        raise ValueError('min_periods must be <= window')
    if center is not False:
        raise ValueError('Method rolling(). The object center\n expected: False')
    if win_type is not None:
        raise ValueError('Method rolling(). The object win_type\n expected: None')
    if on is not None:
        raise ValueError('Method rolling(). The object on\n expected: None')
    if axis != 0:
        raise ValueError('Method rolling(). The object axis\n expected: 0')
    if closed is not None:
        raise ValueError('Method rolling(). The object closed\n expected: None')
    return int(minp)


def _as_kernel_column(values):
    """NumPy column -> contiguous float64 / int64 (the reference reads ints as numbers, output f64)."""
    a = np.ascontiguousarray(values)
    if a.dtype.name in _SUPPORTED_IN:
        return a
    if a.dtype.kind in "iub":
        return a.astype(np.int64)
    if a.dtype.kind == "f":
        return a.astype(np.float64)
    raise TypeError("rolling: unsupported column dtype %s" % a.dtype)


def rolling_host(values, op, window, minp, ddof=1):
    """float64[n] result for a host (NumPy) column, through `sdcb200_host_rolling`."""
    lib = native.load()
    a = _as_kernel_column(values)
    out = np.empty(a.shape[0], dtype=np.float64)
    native.check(lib.sdcb200_host_rolling(
        native.ROLL_OPS[op], native.DTYPE_CODES[a.dtype.name], a.ctypes.data, out.ctypes.data,
        a.shape[0], window, minp, ddof))
    return out


def rolling_device(col, op, window, minp, ddof=1, out=None, halo=None, row_offset=0):
    """Device form: `col` is a 1-D contiguous CUDA torch tensor (float64 / int64).

    `halo` (same dtype) holds the rows that precede col[0] in the global series when the
    column is row-sharded (previous rank's last window-1 rows).  Asynchronous on torch's
    current stream.
    """
    import torch
    lib = native.load()
    assert col.is_cuda and col.dim() == 1 and col.is_contiguous()
    name = {torch.float64: "float64", torch.int64: "int64"}[col.dtype]
    if out is None:
        out = torch.empty(col.shape[0], dtype=torch.float64, device=col.device)
    hptr, hlen = None, 0
    if halo is not None and halo.numel() > 0:
        assert halo.dtype == col.dtype and halo.is_cuda and halo.is_contiguous()
        hptr, hlen = halo.data_ptr(), halo.numel()
    with torch.cuda.device(col.device):
        native.check(lib.sdcb200_rolling(
            native.ROLL_OPS[op], native.DTYPE_CODES[name], col.data_ptr(), out.data_ptr(), col.shape[0],
            window, minp, ddof, hptr, hlen, row_offset, native.current_stream_ptr()))
    return out


class Rolling:
    """`obj.rolling(window, min_periods)` for a pandas Series / DataFrame, NumPy array or CUDA tensor."""

    def __init__(self, obj, window, min_periods=None, center=False, win_type=None, on=None, axis=0, closed=None):
        self._minp = check_rolling_args(window, min_periods, center, win_type, on, axis, closed)
        self._window = int(window)
        self._obj = obj

    def _column(self, values, op, ddof):
        try:
            import torch
            if isinstance(values, torch.Tensor):
                return rolling_device(values, op, self._window, self._minp, ddof)
        except ImportError:
            pass
        return rolling_host(values, op, self._window, self._minp, ddof)

    def _apply(self, op, ddof=1):
        import pandas as pd
        obj = self._obj
        if isinstance(obj, pd.Series):
            return pd.Series(self._column(obj.to_numpy(), op, ddof), index=obj.index, name=obj.name)
        if isinstance(obj, pd.DataFrame):
            cols = {c: self._column(obj[c].to_numpy(), op, ddof) for c in obj.columns}
            return pd.DataFrame(cols, index=obj.index, columns=obj.columns)
        return self._column(obj, op, ddof)

    def sum(self):
        return self._apply("sum")

    def mean(self):
        return self._apply("mean")

    def count(self):
        return self._apply("count")

    def min(self):
        return self._apply("min")

    def max(self):
        return self._apply("max")

    def _check_ddof(self, name, ddof):
        if not isinstance(ddof, (int, np.integer)) or isinstance(ddof, bool):
            raise TypeError('Method rolling.{}(). The object ddof\n given: {}\n expected: int'.format(
                name, type(ddof).__name__))

    def var(self, ddof=1):
        self._check_ddof("var", ddof)
        return self._apply("var", int(ddof))

    def std(self, ddof=1):
        self._check_ddof("std", ddof)
        return self._apply("std", int(ddof))


def rolling(obj, window, min_periods=None, center=False, win_type=None, on=None, axis=0, closed=None):
    return Rolling(obj, window, min_periods, center, win_type, on, axis, closed)
