"""Without a GPU the Numba-bound entry points must compile in nopython mode and FAIL LOUDLY (no CPU fallback)."""
import numba
import numpy as np
import pytest


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="a GPU is present: the call would succeed")
def test_nopython_binding_raises_without_a_device():
    from sdc_b200 import numba_ext as nx

    @numba.njit
    def f(x):
        return nx.rolling_mean(x, 3, 3)

    @numba.njit
    def g(x):
        return nx.parallel_stable_argsort(x, True)

    with pytest.raises(RuntimeError, match="libsdc_b200"):
        f(np.arange(10.0))
    with pytest.raises(RuntimeError, match="libsdc_b200"):
        g(np.arange(10.0))
    # groupby / merge on host arrays type and compile the same way, and fail loudly without a device
    @numba.njit
    def h(k, v):
        return nx.groupby_sum(k, v)

    @numba.njit
    def m(l, r):
        return nx.merge_indexers(l, r, "left")

    with pytest.raises(RuntimeError, match="libsdc_b200"):
        h(np.arange(10), np.arange(10.0))
    with pytest.raises(RuntimeError, match="libsdc_b200"):
        m(np.arange(10), np.arange(10))
    # argument errors are raised in the jitted code before the native call, with the reference's messages
    @numba.njit
    def bad(x):
        return nx.rolling_sum(x, 3, 5)
    with pytest.raises(ValueError, match="min_periods must be <= window"):
        bad(np.arange(10.0))
