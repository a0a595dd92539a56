"""Pin the sort oracle on the reference's own test inputs (pandas / NumPy mergesort as ground truth)."""
from itertools import product

import numpy as np
import pandas as pd

from oracle import sort as osort

ALL_INT = ['int8', 'uint8', 'int16', 'uint16', 'int32', 'uint32', 'int64', 'uint64']
ALL_FLOAT = ['float32', 'float64']


def test_argsort_ascending_matches_numpy_mergesort():
    # sdc/tests/test_sdc_numpy.py:341-386
    rng = np.random.RandomState(0)
    f = [rng.random_sample(10**4), rng.random_sample(10**3)]
    f[1][::2] = np.nan
    ints = [rng.randint(0, 2, 10**4 + 1), np.ones(10**3 + 1, dtype=np.int64), rng.randint(0, 255, 10**3)]
    for case in ALL_FLOAT:
        for a in f:
            d = a.astype(case)
            np.testing.assert_array_equal(osort.stable_argsort(d), np.argsort(d, kind='mergesort'))
    for case in ALL_INT:
        for a in ints:
            d = a.astype(case)
            np.testing.assert_array_equal(osort.stable_argsort(d), np.argsort(d, kind='mergesort'))


def test_argsort_param_ascending_matches_pandas():
    # sdc/tests/test_sdc_numpy.py:388-425
    rng = np.random.RandomState(0)
    values = {'float': [np.inf, -np.inf, np.nan, 0, -1, 2.1, 2 / 3, -3 / 4, 0.777], 'int': [1, -1, 3, 5, -60, 21, 22, 23]}
    dtypes = {'float': ALL_FLOAT, 'int': ALL_INT}
    for ascending in (True, False):
        for group, vals in values.items():
            for dt in dtypes[group]:
                data = rng.choice(vals, 100).astype(dt)
                want = pd.Series(data).sort_values(kind='mergesort', ascending=ascending).index.to_numpy()
                np.testing.assert_array_equal(osort.stable_argsort(data, ascending), want)


def test_negative_zero_ties_keep_input_order():
    a = np.array([0.0, -0.0, 1.0, -0.0, 0.0, np.nan, -1.0])
    np.testing.assert_array_equal(osort.stable_argsort(a, True), [6, 0, 1, 3, 4, 2, 5])
    np.testing.assert_array_equal(osort.stable_argsort(a, False), [2, 0, 1, 3, 4, 6, 5])


def test_sort_values_indexer_matches_pandas():
    # sdc/tests/test_series.py:3856-4072 (float vector with inf / nan / -0.0)
    data = np.array([1, -0., 0.2, -3.7, np.inf, np.nan, -1.0, 2 / 3, 21.2, -np.inf, 9.99])
    s = pd.Series(data)
    for asc, nap in product((True, False), ("last", "first")):
        want = s.sort_values(kind='mergesort', ascending=asc, na_position=nap).index.to_numpy()
        np.testing.assert_array_equal(osort.sort_values_indexer(data, asc, nap), want)
    # examples/series/series_sort_values.py: [3, -10, nan, 0, 92] -> -10, 0, 3, 92, nan
    ex = np.array([3, -10, np.nan, 0, 92])
    np.testing.assert_array_equal(ex[osort.sort_values_indexer(ex)][:4], [-10, 0, 3, 92])


def test_string_argsort_matches_pandas():
    strs = ['b', 'a', '', 'ab', 'a', 'B', 'é', 'zz', 'a\x00', 'abc']
    s = pd.Series(strs)
    for asc in (True, False):
        want = s.sort_values(kind='mergesort', ascending=asc).index.to_numpy()
        np.testing.assert_array_equal(osort.str_stable_argsort(strs, asc), want)
    with_none = ['b', None, 'a', '', None, 'ab']
    for asc, nap in product((True, False), ("last", "first")):
        want = pd.Series(with_none).sort_values(kind='mergesort', ascending=asc, na_position=nap).index.to_numpy()
        np.testing.assert_array_equal(osort.sort_values_indexer(with_none, asc, nap), want)
