"""Pin the restatements against oracle/_ref -- the reference's OWN compiled natives (oracle/ref_recipe builds them with
g++ from /root/reference/sdc/native/str_ext/sdc_str_arr.hpp and sdc/_distributed.h; oracle/ref.py loads them):
string order (stable_argsort), lengths -> offsets, the NA bitmap convention and the block partition."""
import numpy as np
import pytest

from oracle import ref as oref
from oracle import shuffle as oshuffle
from oracle import strings as ostr

pytestmark = pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built (needs /root/reference or the prebuilt file)")


def _random_strings(rng, n, alphabet, max_len):
    alpha = np.array(list(alphabet))
    return ["".join(rng.choice(alpha, rng.integers(0, max_len + 1))) for _ in range(n)]


@pytest.mark.parametrize("asc", [True, False])
def test_string_order_restatement_equals_the_reference(asc):
    rng = np.random.default_rng(5)
    cases = [
        ['ac', 'c', 'cb', 'ca', 'da', 'cc', 'ddd', 'd', ''],                    # sdc/tests/test_series.py:3856-4072 shapes
        ['', 'a', 'ab', 'abc', 'ab', 'a', ''],
        ['b', 'a', 'B', 'A', 'é', 'z', 'Z', 'ß', '日本', 'aa', 'é'],               # bytes above 0x7f compare unsigned
        ['abcdefgh', 'abcdefgha', 'abcdefg', 'abcdefghabcdefgh', 'abcdefghabcdefgz', 'abcdefgh'],
        _random_strings(rng, 5000, "ab", 6),                                       # many ties: stability both ways
        _random_strings(rng, 5000, "abcXYZ019é", 20),
        [],
    ]
    for vals in cases:
        offsets, chars, _ = ostr.to_str_arr(vals)
        np.testing.assert_array_equal(ostr.stable_argsort(offsets, chars, asc), oref.stable_argsort(offsets, chars, asc))


def test_len_arr_to_offsets_equals_the_reference():
    rng = np.random.default_rng(6)
    for n in (0, 1, 7, 1000):
        lens = rng.integers(0, 40, n).astype(np.uint32)
        np.testing.assert_array_equal(ostr.len_arr_to_offsets(lens), oref.len_arr_to_offsets(lens))
    # the shuffle oracle rebuilds offsets the same way (shuffle_utils.py:277-297)
    strs = [[b"ab", b"", b"cde"], [b"f", b"ghij"]]
    dests = [[1, 0, 1], [1, 0]]
    got = oshuffle.shuffle_strings(strs, dests)
    for r in range(2):
        lens = [len(s) for src in range(2) for s, d in zip(strs[src], dests[src]) if d == r]
        np.testing.assert_array_equal(got[r][0], oref.len_arr_to_offsets(np.asarray(lens, dtype=np.uint32)))


def test_na_bitmap_convention_equals_the_reference():
    vals = ['x', None, 'y', None, None, 'z', 'w', 'q', None, 'r']
    _, _, bitmap = ostr.to_str_arr(vals)
    for i, v in enumerate(vals):
        assert oref.is_na(bitmap, i) == (v is None) == ostr.is_na(bitmap, i)


def test_block_partition_equals_the_reference():
    for total in (0, 1, 7, 8, 9, 1000, 1_000_003):
        for P in (1, 2, 3, 8):
            for r in range(P):
                assert oshuffle.block_range(total, r, P) == oref.block_range(total, P, r), (total, P, r)
