"""Pin the join oracle against pandas.merge (the oracle of the reference's own, skipped, join tests)."""
import numpy as np
import pandas as pd
import pytest

from oracle import join as ojoin


def _pandas_pairs(l, r, how):
    L = pd.DataFrame({"k": l, "li": np.arange(len(l), dtype=np.int64)})
    R = pd.DataFrame({"k": r, "ri": np.arange(len(r), dtype=np.int64)})
    m = L.merge(R, on="k", how=how)
    return ojoin.canonical(m["li"].fillna(-1).to_numpy().astype(np.int64), m["ri"].fillna(-1).to_numpy().astype(np.int64))


CASES = {
    # sdc/tests/test_join.py:67-81 (inner, seq), :237-260 (left), :106-120
    "seq1": ([1, 2, 4, 1, 5], [2, 2, 6]) if False else (np.array([3, 5, 1, 0, 4]), np.array([5, 4, 2, 8, 0])),
    "dups": (np.array([3, 1, 1, 3, 4]), np.array([2, 1, 3, 1, 9, 3])),
    "float_nan": (np.array([0.5, np.nan, 2.0, -0.0, np.nan]), np.array([np.nan, 0.0, 2.0, 7.5])),
    "empty_right": (np.array([1, 2, 3]), np.array([], dtype=np.int64)),
    "empty_left": (np.array([], dtype=np.int64), np.array([1, 2, 3])),
}


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_small_cases_match_pandas(case, how):
    l, r = CASES[case]
    got = ojoin.join_indexers(l, r, how)
    want = _pandas_pairs(l, r, how)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
def test_random_match_pandas(how):
    rng = np.random.default_rng(11)
    for nl, nr, hi in ((5000, 700, 1000), (20000, 20000, 3000), (1000, 5000, 100000)):
        l = rng.integers(0, hi, nl)
        r = rng.integers(0, hi, nr)
        got = ojoin.join_indexers(l, r, how)
        want = _pandas_pairs(l, r, how)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])


def test_reference_outer_convention():
    # _sdc_internal_join keeps unmatched rows of BOTH sides with -1 and never matches NaN
    l, r = np.array([1.0, np.nan, 3.0]), np.array([3.0, np.nan, 4.0])
    lidx, ridx = ojoin.outer_indexers(l, r, nan_equal=False)
    pairs = sorted(zip(lidx.tolist(), ridx.tolist()))
    assert pairs == sorted([(0, -1), (2, 0), (-1, 2), (1, -1), (-1, 1)])
