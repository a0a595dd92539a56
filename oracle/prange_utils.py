"""Oracle (test infrastructure): row-range chunking of the reference's prange loops.

Restates `get_chunks` / `parallel_chunks`
(reference: sdc/utilities/prange_utils.py:48-69).  Pinned by the literal
expected lists of sdc/tests/test_prange_utils.py:34-84 (tests/test_oracle_chunks.py).
"""
import numba
import numpy as np


@numba.njit(cache=True)
def chunk_bounds(size, pool_size):
    """Return an int64 array `b` of len(chunks)+1 so chunk i = [b[i], b[i+1]).

    Same split as the reference: `size // pool` rows per chunk and the first
    `size % pool` chunks get one more row (prange_utils.py:55-62); an empty list
    for size < 1 or pool < 1 (:52-53) -> b == [0].
    """
    if size < 1 or pool_size < 1:
        return np.zeros(1, dtype=np.int64)
    parts = min(pool_size, size)
    base = size // parts
    extra = size % parts
    b = np.empty(parts + 1, dtype=np.int64)
    for i in range(parts + 1):
        b[i] = i * base + min(i, extra)
    return b


def get_chunks(size, pool_size):
    """List of (start, stop) tuples, the reference's `Chunk` list."""
    b = chunk_bounds(size, pool_size)
    return [(int(b[i]), int(b[i + 1])) for i in range(len(b) - 1)]


def get_pool_size():
    """reference: prange_utils.py:40-45 with SDC_AUTO_PARALLEL on."""
    return numba.config.NUMBA_NUM_THREADS
