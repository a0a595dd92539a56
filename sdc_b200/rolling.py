"""Series / DataFrame `.rolling(window, min_periods).<op>()` on the B200 backend.

Host-side mirror of the reference's rolling interface:
  * argument checks and messages ........ sdc/datatypes/hpat_pandas_rolling_types.py:108-176
  * Series methods ....................... sdc/datatypes/hpat_pandas_series_rolling_functions.py
        count :886-892, max :1022-1028, mean :1031-1037, min :1049-1055,
        std :1114-1123, sum :1126-1132, var :1135-1144
        skew :1105-1111, kurt :1013-1019, cov :995-1003, corr :797-883
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


def rolling_moments_device(col, op, window, minp, other=None, ddof=1):
    """skew / kurt / cov / corr of 1-D contiguous float64 CUDA tensors (`other` None = the column itself);
    float64[max(len)] on the same device.  Asynchronous on torch's current stream."""
    import torch
    lib = native.load()
    assert col.is_cuda and col.dim() == 1 and col.is_contiguous() and col.dtype == torch.float64
    if other is not None:
        assert other.is_cuda and other.dim() == 1 and other.is_contiguous() and other.dtype == torch.float64
    n_y = col.shape[0] if other is None else other.shape[0]
    out = torch.empty(max(col.shape[0], n_y), dtype=torch.float64, device=col.device)
    with torch.cuda.device(col.device):
        native.check(lib.sdcb200_rolling_moments(
            native.ROLL_MOMENT_OPS[op], col.data_ptr(), col.shape[0], None if other is None else other.data_ptr(), n_y,
            out.data_ptr(), window, minp, ddof, native.current_stream_ptr()))
    return out


def rolling_moments_host(values, op, window, minp, other=None, ddof=1):
    """NumPy in / NumPy out around `rolling_moments_device` (ints are read as numbers, like every rolling op)."""
    import torch
    x = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).cuda()
    y = None if other is None else torch.from_numpy(np.ascontiguousarray(other, dtype=np.float64)).cuda()
    return rolling_moments_device(x, op, window, minp, y, ddof).cpu().numpy()


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

    # ---- skew / kurt / cov / corr (csrc/rolling_moments.cu)
    def _moment_column(self, values, op, other=None, ddof=1):
        try:
            import torch
            if isinstance(values, torch.Tensor):
                return rolling_moments_device(values, op, self._window, self._minp, other, ddof)
        except ImportError:
            pass
        return rolling_moments_host(values, op, self._window, self._minp, other, ddof)

    def _apply_moment(self, op, other=None, ddof=1):
        import pandas as pd
        obj = self._obj
        if isinstance(obj, pd.Series):
            oth = None if other is None else pd.Series(other).to_numpy()
            res = self._moment_column(obj.to_numpy(), op, oth, ddof)
            if other is None:
                return pd.Series(res, index=obj.index, name=obj.name)
            return pd.Series(res)            # two series: the reference returns a default index (:881, :990)
        if isinstance(obj, pd.DataFrame):
            # per-column Series rolling against the same column of `other` (df_rolling_method_other_df_codegen,
            # sdc/datatypes/hpat_pandas_dataframe_rolling_functions.py:83-135) or against the one `other` Series
            cols = {}
            for c in obj.columns:
                oth = None
                if isinstance(other, pd.DataFrame):
                    if c not in other.columns:
                        cols[c] = np.full(len(obj), np.nan)
                        continue
                    oth = other[c].to_numpy()
                elif other is not None:
                    oth = pd.Series(other).to_numpy()
                cols[c] = self._moment_column(obj[c].to_numpy(), op, oth, ddof)
            n = max(len(v) for v in cols.values()) if cols else 0
            cols = {c: (v if len(v) == n else np.concatenate([v, np.full(n - len(v), np.nan)])) for c, v in cols.items()}
            return pd.DataFrame(cols, columns=obj.columns)
        return self._moment_column(obj, op, other, ddof)

    def skew(self):
        return self._apply_moment("skew")

    def kurt(self):
        return self._apply_moment("kurt")

    def _check_other(self, name, other, pairwise):
        import pandas as pd
        ok = other is None or isinstance(other, (pd.Series, pd.DataFrame, np.ndarray)) or hasattr(other, "is_cuda")
        if not ok:
            raise TypeError('Method rolling.{}(). The object other\n given: {}\n expected: Series'.format(
                name, type(other).__name__))
        if pairwise is not None and not isinstance(pairwise, (bool, np.bool_)):
            raise TypeError('Method rolling.{}(). The object pairwise\n given: {}\n expected: bool'.format(
                name, type(pairwise).__name__))

    def cov(self, other=None, pairwise=None, ddof=1):
        self._check_other("cov", other, pairwise)
        self._check_ddof("cov", ddof)
        return self._apply_moment("cov", other, int(ddof))

    def corr(self, other=None, pairwise=None):
        self._check_other("corr", other, pairwise)
        return self._apply_moment("corr", other)


def rolling(obj, window, min_periods=None, center=False, win_type=None, on=None, axis=0, closed=None):
    return Rolling(obj, window, min_periods, center, win_type, on, axis, closed)
