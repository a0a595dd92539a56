"""Oracle (test infrastructure): pd.merge(how='inner' | 'left') on one key column, as row-position pairs.

The reference ships NO merge (sdc/hiframes/join.py:34-43 is a stub, sdc/tests/test_join.py:51-394 is
skipped): PARITY UNPINNED by the reference itself.  What is restated here is the reference's one live
join, the sort-merge OUTER join of two key arrays that Series alignment uses:
    _sdc_internal_join ........ sdc/datatypes/common_functions.py:225-336
        stable argsort of both sides (:255-259), two-pointer walk (:262-309): smaller key -> a row
        with -1 on the other side, equal keys -> the full block cross product in (left, right)
        sorted order, leftovers with -1 (:311-334)
and `how` is applied on top: inner keeps the pairs with both positions, left keeps every pair with a
left position.  One deliberate difference, because the target is pandas.merge (the oracle of the
reference's own skipped join tests): NaN keys match each other (`nan_equal=True`), whereas
_sdc_internal_join never matches NaN (:252-253, :271-279).  Pinned against pandas.merge in
tests/test_oracle_join.py.
"""
import numba
import numpy as np


@numba.njit(cache=True)
def _merge_walk(lk, rk, sl, sr, nan_equal, lidx, ridx, fill):
    """Two-pointer walk over the stable-sorted sides; fills (lidx, ridx) of the OUTER join when
    `fill`, returns the number of result rows either way (the reference grows its buffers
    instead, :242, 266-268)."""
    nl, nr = len(lk), len(rk)
    i = j = k = 0
    while i < nl and j < nr:
        a, b = lk[sl[i]], rk[sr[j]]
        a_nan, b_nan = a != a, b != b
        if (a < b) or (b_nan and not (a_nan and nan_equal)):
            if fill:
                lidx[k] = sl[i]
                ridx[k] = -1
            i += 1
            k += 1
        elif (a > b) or (a_nan and not (b_nan and nan_equal)):
            if fill:
                lidx[k] = -1
                ridx[k] = sr[j]
            j += 1
            k += 1
        else:
            ni, nj = i, j
            while ni < nl and (lk[sl[ni]] == a or (a_nan and lk[sl[ni]] != lk[sl[ni]])):
                ni += 1
            while nj < nr and (rk[sr[nj]] == b or (b_nan and rk[sr[nj]] != rk[sr[nj]])):
                nj += 1
            for s in range(i, ni):
                for t in range(j, nj):
                    if fill:
                        lidx[k] = sl[s]
                        ridx[k] = sr[t]
                    k += 1
            i, j = ni, nj
    while i < nl:
        if fill:
            lidx[k] = sl[i]
            ridx[k] = -1
        i += 1
        k += 1
    while j < nr:
        if fill:
            lidx[k] = -1
            ridx[k] = sr[j]
        j += 1
        k += 1
    return k


def outer_indexers(left_keys, right_keys, nan_equal=True):
    lk = np.ascontiguousarray(left_keys)
    rk = np.ascontiguousarray(right_keys)
    if lk.dtype != rk.dtype:
        lk, rk = lk.astype(np.float64), rk.astype(np.float64)
    sl = np.argsort(lk, kind="stable").astype(np.int64)      # NaN last, like the reference's argsort
    sr = np.argsort(rk, kind="stable").astype(np.int64)
    dummy = np.empty(0, np.int64)
    k = _merge_walk(lk, rk, sl, sr, nan_equal, dummy, dummy, False)
    lidx, ridx = np.empty(k, np.int64), np.empty(k, np.int64)
    _merge_walk(lk, rk, sl, sr, nan_equal, lidx, ridx, True)
    return lidx, ridx


def join_indexers(left_keys, right_keys, how="inner"):
    """-> (lidx, ridx) int64 in canonical order (sorted by left row, then right row)."""
    assert how in ("inner", "left", "right", "outer")
    lidx, ridx = outer_indexers(left_keys, right_keys, True)
    keep = {"inner": (lidx >= 0) & (ridx >= 0), "left": lidx >= 0, "right": ridx >= 0,
            "outer": np.ones(len(lidx), dtype=bool)}[how]
    lidx, ridx = lidx[keep], ridx[keep]
    order = np.lexsort((ridx, lidx))
    return lidx[order], ridx[order]


def canonical(lidx, ridx):
    """Sort (lidx, ridx) pairs so two joins can be compared as row sets."""
    lidx, ridx = np.asarray(lidx), np.asarray(ridx)
    order = np.lexsort((ridx, lidx))
    return lidx[order], ridx[order]
