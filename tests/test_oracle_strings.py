"""Pin the string oracle: layout round trip and ordering against pandas on the reference's test inputs."""
import numpy as np
import pandas as pd

from oracle import strings as ostr

# sdc/tests/test_series.py:3856-4072 string cases (None and '' included) + prefix / non-ASCII cases
CASES = [
    ['a', 'd', 'r', 'cc'],
    ['ac', 'c', 'cb', 'ca', None, 'da', 'cc', 'ddd', 'd', ''],
    ['', 'a', 'ab', 'abc', 'ab', 'a', ''],
    ['b', 'a', 'B', 'A', 'é', 'z', 'Z', 'ß', '日本', 'aa'],
    ['abcdefgh', 'abcdefgha', 'abcdefg', 'abcdefghabcdefgh', 'abcdefghabcdefgz', 'abcdefgh'],
]


def test_layout_round_trip():
    vals = CASES[1]
    offsets, chars, bitmap = ostr.to_str_arr(vals)
    assert offsets[0] == 0 and offsets[-1] == len(chars)
    for i, v in enumerate(vals):
        assert ostr.is_na(bitmap, i) == (v is None)
        assert chars[offsets[i]:offsets[i + 1]].tobytes() == (b"" if v is None else v.encode())


def test_order_matches_pandas_both_directions():
    for vals in CASES:
        nn = [v for v in vals if v is not None]
        offsets, chars, _ = ostr.to_str_arr(nn)
        s = pd.Series(nn, dtype=object)
        for asc in (True, False):
            got = ostr.stable_argsort(offsets, chars, asc)
            want = s.sort_values(kind="mergesort", ascending=asc).index.to_numpy()
            np.testing.assert_array_equal(got, want, err_msg="%s asc=%s" % (vals, asc))


def test_random_against_python_sort():
    rng = np.random.default_rng(0)
    alphabet = np.array(list("abcXYZ019"))
    vals = ["".join(rng.choice(alphabet, rng.integers(0, 12))) for _ in range(3000)]
    offsets, chars, _ = ostr.to_str_arr(vals)
    got = ostr.stable_argsort(offsets, chars, True)
    want = np.asarray(sorted(range(len(vals)), key=lambda i: vals[i].encode()), dtype=np.uint64)
    np.testing.assert_array_equal(got, want)
