"""ORACLE -- test infrastructure only (never imported by sdc_b200/ or bench.py's GPU arm).

CPU restatement of the reference's CSV ingest for numeric tables: `pd.read_csv` lowered to Arrow's reader
(sdc/io/csv_ext.py:93-110 do_read_csv; options :174-272).  Arrow itself is importable here, so the oracle IS the
reference's call -- `pyarrow.csv.read_csv` with the options csv_ext.py builds -- followed by the reference's boxing
rule (nulls become NaN, which makes an integer column float64: ArrowTableType -> pandas, csv_ext.py:112-141).
`parse_float_field` is the independent second opinion for single fields: Python's float() is correctly rounded.
"""
import io

import numpy as np


def read_csv_numeric(data, delimiter=",", names=None, usecols=None, dtype=None, skiprows=None):
    """data: bytes -> {name: np.ndarray int64 / float64}, as the reference would hand the columns to NumPy."""
    import pyarrow as pa
    import pyarrow.csv as pacsv
    # csv_ext.py:174-189: names=None -> the first row holds the names; names given -> autogenerate=False, every row is data
    read_opts = pacsv.ReadOptions(skip_rows=int(skiprows or 0), column_names=list(names) if names is not None else None)
    parse_opts = pacsv.ParseOptions(delimiter=delimiter)                                  # csv_ext.py:191-197
    column_types = None
    if dtype is not None:                                                                 # csv_ext.py:217-231
        if isinstance(dtype, dict):
            column_types = {k: pa.from_numpy_dtype(np.dtype(v)) for k, v in dtype.items()}
    include = None
    convert_opts = pacsv.ConvertOptions(column_types=column_types, include_columns=include, strings_can_be_null=True)
    table = pacsv.read_csv(io.BytesIO(data), read_options=read_opts, parse_options=parse_opts, convert_options=convert_opts)
    all_names = table.column_names
    if usecols is not None:
        keep = sorted({(all_names.index(u) if isinstance(u, str) else int(u)) for u in usecols})
    else:
        keep = list(range(len(all_names)))
    out = {}
    for i in keep:
        col = table.column(i)
        want = None if dtype is None or isinstance(dtype, dict) else np.dtype(dtype)
        if want is not None:
            col = col.cast(pa.from_numpy_dtype(want))
        if col.null_count:
            col = col.cast(pa.float64())
            a = col.to_numpy()
            a = np.where(np.asarray(col.is_null()), np.nan, a)
        else:
            a = col.to_numpy()
        out[all_names[i]] = np.ascontiguousarray(a)
    return out


def parse_float_field(text):
    return float(text)
