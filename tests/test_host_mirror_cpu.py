"""CPU checks of the host-side mirror (sdc_b200/*.py): the argument validation and the error messages the reference
raises BEFORE any native call (so they need no GPU), and that there is no CPU fallback behind them -- a valid call
without a GPU fails loudly instead of computing something on the host.

Messages follow sdc/datatypes/hpat_pandas_rolling_types.py:144-170 (pinned by sdc/tests/test_rolling.py:426-475),
hpat_pandas_series_functions.py:3921-3933 (sort_values), hpat_pandas_dataframe_functions.py:3066-3067 and
hpat_pandas_groupby_functions.py:169-170 (groupby), hpat_pandas_series_functions.py:4797-4798 (Series.groupby).
"""
import numpy as np
import pandas as pd
import pytest
import torch

from sdc_b200 import groupby as gb
from sdc_b200 import join as jn
from sdc_b200 import rolling as ro
from sdc_b200 import sort as so


def test_rolling_argument_errors_match_the_reference():
    s = pd.Series([1.0, 2.0, 3.0])
    with pytest.raises(ValueError, match="window must be non-negative"):
        ro.rolling(s, -1)
    with pytest.raises(ValueError, match="min_periods must be >= 0"):
        ro.rolling(s, 3, -1)
    with pytest.raises(ValueError, match="min_periods must be <= window"):
        ro.rolling(s, 3, 4)
    with pytest.raises(TypeError, match="The object window"):
        ro.rolling(s, 2.5)
    with pytest.raises(TypeError, match="The object min_periods"):
        ro.rolling(s, 3, "2")
    for kw, bad in (("center", True), ("win_type", "triang"), ("on", "A"), ("axis", 1), ("closed", "right")):
        with pytest.raises(ValueError, match="The object " + kw):
            ro.rolling(s, 3, 2, **{kw: bad})
    assert ro.check_rolling_args(5) == 5 and ro.check_rolling_args(5, 0) == 0          # min_periods=None -> window


def test_sort_values_argument_errors_match_the_reference():
    s = pd.Series([3.0, 1.0, 2.0])
    with pytest.raises(ValueError, match="kind != 'quicksort', 'mergesort'"):
        so.series_sort_values(s, kind="heapsort")
    with pytest.raises(TypeError, match="Given inplace"):
        so.series_sort_values(s, inplace=True)
    with pytest.raises(ValueError, match="na_position != 'last', 'first'"):
        so.series_sort_values(s, na_position="middle")
    with pytest.raises(ValueError, match="The object axis"):
        so.series_sort_values(s, axis=1)
    with pytest.raises(ValueError, match="Unsupported value of 'kind' parameter"):
        so.argsort(np.arange(3), kind="heapsort")


def test_groupby_argument_errors_match_the_reference():
    df = pd.DataFrame({"A": [1, 2, 1], "B": [1.0, 2.0, 3.0], "C": [4.0, 5.0, 6.0]})
    with pytest.raises(TypeError, match="The object by"):
        gb.groupby(df, ["A", "B"])                      # one literal column only (:3066-3067)
    with pytest.raises(TypeError, match="The object by"):
        gb.groupby(df, "Z")
    sel = gb.groupby(df, "A")["B", "C"]
    with pytest.raises(IndexError, match="Columns already selected"):
        sel["B"]                                        # second [] (:169-170)
    with pytest.raises(KeyError):
        gb.groupby(df, "A")["nope"]
    with pytest.raises(ValueError, match="array of the same length"):
        gb.groupby(df["B"], np.arange(2))               # :4797-4798
    with pytest.raises(TypeError, match="The object ddof"):
        gb.groupby(df, "A").std(ddof=1.5)
    with pytest.raises(TypeError, match="expected a pandas DataFrame or Series"):
        gb.groupby([1, 2, 3], "A")


def test_merge_argument_errors():
    l = pd.DataFrame({"k": [1, 2, 3], "x": [1.0, 2.0, 3.0]})
    r = pd.DataFrame({"k": [1, 2, 4], "y": [4.0, 5.0, 6.0]})
    with pytest.raises(ValueError, match="is not supported"):
        jn.merge(l, r, how="cross", on="k")
    with pytest.raises(ValueError, match='Can only pass argument "on" OR "left_on" and "right_on"'):
        jn.merge(l, r, on="k", left_on="k")             # pandas' own message
    with pytest.raises(TypeError, match="exactly one key column"):
        jn.merge(l, r, on=["k", "x"])
    with pytest.raises(ValueError, match="exactly one key column"):
        jn.merge(l, pd.DataFrame({"z": [1]}))           # no common column to infer the key from
    with pytest.raises(TypeError, match="key dtype"):
        jn.merge(pd.DataFrame({"k": ["a", "b"]}), pd.DataFrame({"k": ["a", "c"]}), on="k")


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_no_cpu_fallback_behind_the_mirror():
    """A valid call on a box without a GPU raises (the CUDA runtime / torch refuses): nothing is computed on the host."""
    s = pd.Series(np.arange(10, dtype=np.float64))
    with pytest.raises(Exception):
        ro.rolling(s, 3).mean()
    with pytest.raises(Exception):
        gb.groupby(pd.DataFrame({"A": [1, 2, 1], "B": [1.0, 2.0, 3.0]}), "A").sum()
    with pytest.raises(Exception):
        so.argsort(np.array([3, 1, 2]))
