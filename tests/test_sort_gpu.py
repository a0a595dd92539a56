"""GPU parity: csrc/sort.cu (onesweep radix sort / argsort / gather) against the oracle -- index-exact."""
from itertools import product

import numpy as np
import pandas as pd
import pytest

from oracle import sort as osort

pytestmark = pytest.mark.gpu

ALL_INT = ['int8', 'uint8', 'int16', 'uint16', 'int32', 'uint32', 'int64', 'uint64']
ALL_FLOAT = ['float32', 'float64']


def test_sort_all_dtypes_reference_case():
    # sdc/tests/test_sdc_numpy.py:311-339
    from sdc_b200.sort import sort
    rng = np.random.RandomState(0)
    fa, ia = rng.random_sample(100), rng.randint(0, 127, 100)
    for kind in (None, 'quicksort', 'mergesort'):
        for dt in ALL_FLOAT:
            a = fa.astype(dt)
            np.testing.assert_array_equal(sort(a.copy(), kind), np.sort(a, kind='stable'))
        for dt in ALL_INT:
            a = ia.astype(dt)
            np.testing.assert_array_equal(sort(a.copy(), kind), np.sort(a, kind='stable'))


def test_argsort_reference_cases_index_exact():
    # sdc/tests/test_sdc_numpy.py:341-386 (sizes up to 10**6 + 1, NaN in every second element)
    from sdc_b200.sort import argsort
    rng = np.random.RandomState(0)
    f = [rng.random_sample(10**5), rng.random_sample(10**4)]
    f[1][::2] = np.nan
    ints = [rng.randint(0, 2, 10**6 + 1), np.ones(10**5 + 1, dtype=np.int64), rng.randint(0, 255, 10**4)]
    for dt in ALL_FLOAT:
        for a in f:
            d = a.astype(dt)
            np.testing.assert_array_equal(argsort(d, 'mergesort'), osort.stable_argsort(d))
    for dt in ALL_INT:
        for a in ints:
            d = a.astype(dt)
            np.testing.assert_array_equal(argsort(d, 'mergesort'), osort.stable_argsort(d))


def test_argsort_param_ascending():
    # sdc/tests/test_sdc_numpy.py:388-425
    from sdc_b200.sort import argsort
    rng = np.random.RandomState(0)
    values = {'float': [np.inf, -np.inf, np.nan, 0, -1, 2.1, 2 / 3, -3 / 4, 0.777], 'int': [1, -1, 3, 5, -60, 21, 22, 23]}
    dtypes = {'float': ALL_FLOAT, 'int': ALL_INT}
    for ascending in (True, False):
        for group, vals in values.items():
            for dt in dtypes[group]:
                data = rng.choice(vals, 100).astype(dt)
                np.testing.assert_array_equal(argsort(data, 'mergesort', ascending),
                                              osort.stable_argsort(data, ascending), err_msg="%s %s" % (dt, ascending))


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 3071, 3072, 3073, 4096, 50_000, 1_000_003])
def test_device_argsort_sizes_and_ranges(n):
    import torch
    from sdc_b200.sort import argsort_device, gather_device, sort_device
    rng = np.random.default_rng(n)
    cases = {
        "i64_full": rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64),
        "i64_small": rng.integers(-5, 5, n, dtype=np.int64),
        "u64": rng.integers(0, 2**64 - 1, n, dtype=np.uint64),
        "i32": rng.integers(-2**31, 2**31 - 1, n, dtype=np.int32),
        "u16": rng.integers(0, 2**16 - 1, n, dtype=np.uint16),
        "i8": rng.integers(-128, 127, n, dtype=np.int8),
        "f32": (rng.standard_normal(n) * 1e3).astype(np.float32),
    }
    f = rng.standard_normal(n) * 1e6
    if n > 10:
        f[rng.integers(0, n, n // 10)] = np.nan
        f[rng.integers(0, n, n // 20)] = 0.0
        f[rng.integers(0, n, n // 20)] = -0.0
        f[rng.integers(0, n, 3)] = np.inf
        f[rng.integers(0, n, 3)] = -np.inf
    cases["f64"] = f
    for name, a in cases.items():
        if a.dtype in (np.uint64, np.uint16):
            d = torch.from_numpy(a.view(a.dtype.str.replace('u', 'i'))).cuda().view(
                {np.dtype('uint64'): torch.uint64, np.dtype('uint16'): torch.uint16}[a.dtype])
        else:
            d = torch.from_numpy(a).cuda()
        for asc in (True, False):
            got = argsort_device(d, asc).cpu().numpy()
            np.testing.assert_array_equal(got, osort.stable_argsort(a, asc), err_msg="%s asc=%s n=%d" % (name, asc, n))
        if n and a.dtype.itemsize in (1, 2, 4, 8) and a.dtype not in (np.uint64, np.uint16):
            idx = argsort_device(d, True)
            np.testing.assert_array_equal(gather_device(d, idx).cpu().numpy(), a[osort.stable_argsort(a)])
        if a.dtype not in (np.uint64, np.uint16):
            s = sort_device(d.clone()).cpu().numpy()
            np.testing.assert_array_equal(s, np.sort(a, kind='stable'))


def test_keys_only_stable_sort_keeps_signed_zero_order():
    from sdc_b200.sort import sort
    a = np.array([0.0, -0.0, 1.0, -0.0, 0.0, np.nan, -1.0] * 1000)
    got = sort(a.copy())
    want = a[osort.stable_argsort(a)]
    assert np.array_equal(np.signbit(got), np.signbit(want))
    np.testing.assert_array_equal(got, want)


def test_reference_named_symbols():
    """parallel_stable_argsort_u64f64 & co. bound like sdc/functions/sort.py:39-72 does."""
    import ctypes as ct
    from sdc_b200 import native
    lib = native.load()
    a = np.random.default_rng(0).random(10_000)
    a[::7] = np.nan
    for asc in (1, 0):
        idx = np.empty(a.size, dtype=np.uint64)
        lib.parallel_stable_argsort_u64f64(idx.ctypes.data, a.ctypes.data, a.size, asc)
        np.testing.assert_array_equal(idx.astype(np.int64), osort.stable_argsort(a, bool(asc)))
    b = np.random.default_rng(1).integers(-100, 100, 5000).astype(np.int16)
    c = b.copy()
    lib.parallel_sort_i16(c.ctypes.data, c.size)
    np.testing.assert_array_equal(c, np.sort(b))
    lib.set_number_of_threads(8)


def test_series_sort_values_matches_pandas():
    # sdc/tests/test_series.py:3856-4072
    from sdc_b200.sort import series_sort_values
    data = [1, -0., 0.2, -3.7, np.inf, np.nan, -1.0, 2 / 3, 21.2, -np.inf, 9.99]
    s = pd.Series(data, index=np.arange(len(data))[::-1], name="A")
    for asc, nap in product((True, False), ("last", "first")):
        got = series_sort_values(s, ascending=asc, kind='mergesort', na_position=nap)
        pd.testing.assert_series_equal(got, s.sort_values(ascending=asc, kind='mergesort', na_position=nap))
    ex = pd.Series([3, -10, np.nan, 0, 92])          # examples/series/series_sort_values.py
    pd.testing.assert_series_equal(series_sort_values(ex), ex.sort_values(kind='mergesort'))
    with pytest.raises(ValueError, match="kind"):
        series_sort_values(s, kind='heapsort')
    with pytest.raises(ValueError, match="na_position"):
        series_sort_values(s, na_position='middle')


def test_large_sortedness_and_permutation_property():
    """C4-scale numeric keys (200 M rows would need 6.4 GB; 50 M here): sorted + a permutation."""
    import torch
    from sdc_b200.sort import argsort_device, gather_device
    n = 50_000_000
    g = torch.Generator(device="cuda").manual_seed(4)
    d = torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g)
    idx = argsort_device(d, True)
    s = gather_device(d, idx)
    assert bool((s[1:] >= s[:-1]).all())
    assert int(torch.bincount(idx, minlength=n).max().item()) == 1 and int(idx.min()) == 0 and int(idx.max()) == n - 1
    assert int(s.sum().item()) == int(d.sum().item())
