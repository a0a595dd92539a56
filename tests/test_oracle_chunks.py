"""Oracle pin: get_chunks against the literal lists of sdc/tests/test_prange_utils.py:34-84."""
from oracle.prange_utils import get_chunks

KNOWN = [
    ((5, 5), [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]),
    ((5, 6), [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]),
    ((9, 5), [(0, 2), (2, 4), (4, 6), (6, 8), (8, 9)]),
    ((9, 4), [(0, 3), (3, 5), (5, 7), (7, 9)]),
    ((9, 2), [(0, 5), (5, 9)]),
    ((9, 3), [(0, 3), (3, 6), (6, 9)]),
    ((0, 5), []),
    ((5, 0), []),
]


def test_get_chunks_known_answers():
    for args, want in KNOWN:
        assert get_chunks(*args) == want, args
