import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
sys.path.insert(0, GOLDEN)
import inputs as golden_inputs  # noqa: E402,F401


def rolling_golden_cases():
    """Yield (op, data, window, minp, ddof, expected) from tests/golden/rolling_pandas.npz."""
    z = np.load(os.path.join(GOLDEN, "rolling_pandas.npz"))
    for key in z.files:
        op, dname, w, mp, ddof = key.split("|")
        src = golden_inputs.GLOBAL_FLOAT64 if dname[0] == "g" else golden_inputs.ROLLING_SUM_DATA
        data = np.asarray(src[int(dname[1:])], dtype=np.float64)
        yield op, data, int(w), int(mp), int(ddof), z[key]


def assert_close(got, want, rtol=1e-9, atol=1e-12, msg=""):
    """NaN pattern must match exactly; finite values within rtol/atol; infs equal."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, "%s shape %s vs %s" % (msg, got.shape, want.shape)
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), "%s NaN pattern differs\n got %s\nwant %s" % (msg, got, want)
    ok = ~gn
    if ok.any():
        np.testing.assert_allclose(got[ok], want[ok], rtol=rtol, atol=atol, err_msg=msg)
