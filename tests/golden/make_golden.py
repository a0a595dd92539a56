"""Generate tests/golden/*.npz with pandas (the reference tests' own oracle) on the
reference's test inputs (tests/golden/inputs.py).  Run here, in the authoring container:

    python tests/golden/make_golden.py

The sweeps follow the reference helpers: sdc/tests/test_rolling.py:619-631 (count),
:696-709 / :742-755 (max / min), :712-724 (mean), :843-855 (sum), :820-831 / :858-868 (std / var).
"""
import os
import sys
from itertools import product

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import inputs  # noqa: E402


def rolling_cases():
    out = {}
    for di, data in enumerate(inputs.GLOBAL_FLOAT64):
        s = pd.Series(data, dtype="float64")
        n = len(data)
        for w in range(0, n + 3, 2):
            for mp in range(0, w + 1, 2):
                out["count|g%d|%d|%d|1" % (di, w, mp)] = s.rolling(w, mp).count().to_numpy()
        for op in ("min", "max"):
            for w in range(1, n + 2):
                for mp in range(w + 1):
                    out["%s|g%d|%d|%d|1" % (op, di, w, mp)] = getattr(s.rolling(w, mp), op)().to_numpy()
    for di, data in enumerate(inputs.ROLLING_SUM_DATA):
        s = pd.Series(data, dtype="float64")
        n = len(data)
        for op in ("sum", "mean"):
            for w in range(n + 2):
                for mp in range(w):
                    out["%s|s%d|%d|%d|1" % (op, di, w, mp)] = getattr(s.rolling(w, mp), op)().to_numpy()
        for op in ("std", "var"):
            for w in range(0, n + 3, 2):
                for mp, ddof in product(range(0, w, 2), [0, 1]):
                    out["%s|s%d|%d|%d|%d" % (op, di, w, mp, ddof)] = getattr(s.rolling(w, mp), op)(ddof).to_numpy()
    return out


def main():
    cases = rolling_cases()
    np.savez_compressed(os.path.join(HERE, "rolling_pandas.npz"), **cases)
    print("rolling cases:", len(cases), "pandas", pd.__version__)


if __name__ == "__main__":
    main()
