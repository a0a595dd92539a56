"""Inputs of the reference's own tests for the hot path, restated (values only).

rolling: sdc/tests/test_rolling.py:1606-1610, 1672-1702 and sdc/tests/test_utils.py:53-62
examples: the literal expected outputs in examples/series/rolling/*.py,
          examples/dataframe/rolling/dataframe_rolling_mean.py,
          examples/dataframe/groupby/*.py, examples/series/series_groupby.py
"""
import numpy as np

MIN64 = np.finfo('float64').min
MAX64 = np.finfo('float64').max
NAN, INF = np.nan, np.inf

# test_utils.py:56-62 (test_global_input_data_float64)
GLOBAL_FLOAT64 = [
    [1., -1., 0.1, MIN64, MAX64, MAX64, MIN64, -0.1],
    [1., NAN, -1., 0., MIN64, MAX64, MAX64, MIN64],
    [1., INF, INF, -1., 0., INF, -INF, -INF],
    [NAN, INF, INF, NAN, NAN, NAN, -INF, -0.0],
]

# test_rolling.py:1672-1688 (mean / sum / std / var series tests)
ROLLING_SUM_DATA = [
    list(range(10)), [1., -1., 0., 0.1, -0.1],
    [1., INF, INF, -1., 0., INF, -INF, -INF],
    [NAN, INF, INF, NAN, NAN, NAN, -INF, -0.0],
]

# literal known-answer vectors from the example scripts' comments
EXAMPLES_ROLLING = {
    # op: (input, window, expected printed to 6 decimals)
    "sum": ([4, 3, 5, 2, 6], 3, [NAN, NAN, 12.0, 10.0, 13.0]),
    "mean": ([4, 3, 5, 2, 6], 3, [NAN, NAN, 4.0, 3.333333, 4.333333]),
    "std": ([4, 3, 5, 2, 6], 3, [NAN, NAN, 1.0, 1.527525, 2.081666]),
    "var": ([4, 3, 5, 2, 6], 3, [NAN, NAN, 1.0, 2.333333, 4.333333]),
    "min": ([4, 3, 5, 2, 6], 3, [NAN, NAN, 3.0, 2.0, 2.0]),
    "max": ([4, 3, 5, 2, 6], 3, [NAN, NAN, 5.0, 5.0, 6.0]),
    "count": ([4, 3, 2, NAN, 6], 3, [1.0, 2.0, 3.0, 2.0, 2.0]),
}
EXAMPLE_DF_ROLLING_MEAN = {
    "input": {"A": [4, 3, 5, 2, 6], "B": [-4, -3, -5, -2, -6]}, "window": 3,
    "expected": {"A": [NAN, NAN, 4.0, 3.333333, 4.333333], "B": [NAN, NAN, -4.0, -3.333333, -4.333333]},
}

EXAMPLE_GROUPBY_DF = {"A": [1, 2, 3, 1, 2, 3, 3, 3, 2],
                      "B": [0, 1, 5, 0, 2, 4, 3, 2, 3],
                      "C": [1, 2, 3, 4, 5, 6, 7, 8, 9]}
EXAMPLE_GROUPBY_COUNT_DF = {"A": [1, 2, 3, 1, 2, 3, 3, 3, 2],
                            "B": [0, 1, NAN, NAN, 2, 4, 3, 2, INF],
                            "C": [NAN, 2, 3, NAN, 5, 6, 7, 8, 9]}
EXAMPLES_GROUPBY = {   # index is [1, 2, 3] for all
    "sum": {"B": [0, 6, 14], "C": [5, 16, 24]},
    "mean": {"B": [0.0, 2.0, 3.5], "C": [2.5, 5.333333, 6.0]},
    "min": {"B": [0, 1, 2], "C": [1, 2, 3]},
    "max": {"B": [0, 3, 5], "C": [4, 9, 8]},
    "std": {"B": [0.0, 1.0, 1.290994], "C": [2.121320, 3.511885, 2.160247]},
    "var": {"B": [0.0, 1.0, 1.666667], "C": [4.5, 12.333333, 4.666667]},
    "count": {"B": [1, 3, 3], "C": [0, 3, 4]},    # on EXAMPLE_GROUPBY_COUNT_DF
}
EXAMPLE_SERIES_GROUPBY = {"S": [390., 350., 30., 20.], "by": [0, 1, 0, 1], "mean": [210.0, 185.0], "index": [0, 1]}

# sdc/tests/test_groupby.py:58-64 (_default_df_numeric_data) and :83-100 (key dtypes)
GROUPBY_DEFAULT_DF = {
    'A': [2, 1, 2, 1, 2, 2, 1, 0, 3, 1, 3],
    'B': list(range(11)),
    'C': [float(i) for i in range(11)],
    'D': [NAN, 2., -1.3, NAN, 3.5, 0, 10, 0.42, NAN, -2.5, 23],
    'E': [INF, 2., -1.3, -INF, 3.5, 0, 10, 0.42, NAN, -2.5, 23],
}
GROUPBY_KEY_DTYPES = {
    'int': [2, 1, 1, 1, 2, 2, 1, 0, 3, 1, 3],
    'float': [2, 1, 1, 1, 2, 2, 1, 3, NAN, 1, NAN],
    'string': ['b', 'a', 'a', 'a', 'b', 'b', 'a', ' ', None, 'a', None],
}
