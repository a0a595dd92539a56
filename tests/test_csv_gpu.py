"""GPU parity: csrc/csv.cu (CSV text -> int64 / float64 columns on the device) against the reference's own reader call
(oracle/csv.py = pyarrow.csv.read_csv with the options sdc/io/csv_ext.py builds) -- bit-exact."""
import io
import struct

import numpy as np
import pytest

from oracle import csv as ocsv

pytestmark = pytest.mark.gpu


def _check(data, **kw):
    from sdc_b200.csv import read_csv_device
    want = ocsv.read_csv_numeric(data, delimiter=kw.get("delimiter", kw.get("sep", ",")), names=kw.get("names"),
                                 usecols=kw.get("usecols"), dtype=kw.get("dtype"), skiprows=kw.get("skiprows"))
    cols, n = read_csv_device(data, **kw)
    assert list(cols) == list(want), (list(cols), list(want))
    for name, w in want.items():
        g = cols[name].cpu().numpy()
        assert g.dtype == w.dtype, (name, g.dtype, w.dtype)
        assert g.shape == w.shape, (name, g.shape, w.shape)
        if w.dtype.kind == "f":
            # NaN payloads aside, the bits must agree
            gn, wn = np.isnan(g), np.isnan(w)
            np.testing.assert_array_equal(gn, wn, err_msg=name)
            np.testing.assert_array_equal(g[~gn].view(np.uint64), w[~wn].view(np.uint64), err_msg=name)
        else:
            np.testing.assert_array_equal(g, w, err_msg=name)
    return cols, n


def test_small_literals_header_crlf_blank_lines_nulls():
    _check(b"a,b,c\n1,1.5,\n2,,3\n\n4,nan,5\r\n")
    _check(b"a,b\r\n1,2\r\n3,4")                                # CRLF, no trailing newline
    _check(b"\n\na,b\n\n1,2\n\n\n3,4\n\n")                      # empty lines anywhere
    _check(b"x\n-0.0\n+1.5\n 2.5 \n.5\n5.\n1e400\n-1e400\n1e-400\nINF\n-inf\nNAN\n")
    _check(b"x\n9223372036854775807\n-9223372036854775808\n012\n 7\n")
    _check(b"x\n9223372036854775808\n1\n")                      # one field beyond int64: the column is float64
    _check(b"i,f\n1,NA\n2,N/A\n3,null\n4,#N/A\n5,1.#QNAN\n")
    _check(b"1;2.5\n3;4.5\n", delimiter=";", names=["p", "q"])   # names given: every line is data
    _check(b"junk line\nmore junk\na,b\n1,2\n", skiprows=2)
    _check(b"a,b,c\n1,2,3\n4,5,6\n", usecols=["c", "a"])
    _check(b"a,b,c\n1,2,3\n4,5,6\n", usecols=[1])
    _check(b"a,b\n1,2\n3,4\n", dtype={"a": np.float64})
    _check(b"a,b\n1,2\n3,4\n", dtype=np.float64)
    _check(b"a\n1")                                            # one field, no newline


@pytest.mark.parametrize("shape", ["narrow", "wide", "ragged_widths", "long_rows"])
def test_random_tables_bit_exact(shape):
    """Random numeric tables written with several float notations: repr (17 digits), fixed, scientific, integers,
    missing fields; narrow rows (256 per CTA staged in shared memory), census-shaped wide rows (fewer rows per CTA) and
    rows too long to stage (read from global memory)."""
    rng = np.random.default_rng({"narrow": 1, "wide": 2, "ragged_widths": 3, "long_rows": 4}[shape])
    nrows = {"narrow": 200_003, "wide": 40_001, "ragged_widths": 50_000, "long_rows": 3_000}[shape]
    ncols = {"narrow": 3, "wide": 24, "ragged_widths": 6, "long_rows": 300}[shape]
    cols = []
    for c in range(ncols):
        kind = c % 6
        if kind == 0:
            v = [repr(float(x)) for x in rng.standard_normal(nrows) * 10.0 ** rng.integers(-8, 8)]
        elif kind == 1:
            v = [str(int(x)) for x in rng.integers(-2**62, 2**62, nrows)]
        elif kind == 2:
            v = ["%.6f" % x for x in rng.uniform(-1e6, 1e6, nrows)]
        elif kind == 3:
            v = ["%.10e" % x for x in rng.standard_normal(nrows) * 1e200]
        elif kind == 4:
            bits = rng.integers(0, 2**63, nrows, dtype=np.uint64)       # any finite double, shortest repr
            xs = bits.view(np.float64)
            v = [repr(float(x)) if np.isfinite(x) else "0.5" for x in xs]
        else:
            v = [str(int(x)) if rng.random() > 0.1 else "" for x in rng.integers(-1000, 1000, nrows)]
        if shape == "ragged_widths" and c == 0:
            v = [s if i % 7 else s + "0" * (i % 300) for i, s in enumerate(v)]       # a few very wide fields
        cols.append(v)
    names = ["c%d" % c for c in range(ncols)]
    lines = [",".join(names)] + [",".join(row) for row in zip(*cols)]
    data = ("\n".join(lines) + "\n").encode()
    got, n = _check(data)
    assert n == nrows


def test_long_digit_strings_on_the_device():
    """Fields with more than 19 significant digits, ties and near-ties included: the device's big-integer tie-break
    (num_parse.cuh round_long_decimal, local memory) gives Arrow's bits."""
    from fractions import Fraction
    import random
    rng = random.Random(5)
    fields = []
    for _ in range(3000):
        m = rng.getrandbits(53) | (1 << 52)
        e = rng.randint(-1000, 900)
        v = Fraction(2 * m + 1, 2) * (Fraction(2) ** e)
        k2 = v.denominator.bit_length() - 1
        digits = str(v.numerator * 5 ** k2)
        fields.append(digits + ("e-%d" % k2 if k2 else ""))
        fields.append(digits + "0000001" + "e-%d" % (k2 + 7))
        fields.append("%.30e" % struct.unpack("<d", struct.pack("<Q", rng.getrandbits(62)))[0])
    data = ("x\n" + "\n".join(fields) + "\n").encode()
    _check(data)


def test_errors_are_loud():
    from sdc_b200.csv import NotNumericCSV, read_csv_device
    with pytest.raises(NotNumericCSV):
        read_csv_device(b"a,b\n1,x\n2,y\n")
    with pytest.raises(NotNumericCSV):
        read_csv_device(b'a,b\n1,"2"\n')
    with pytest.raises(RuntimeError, match="row 1, column 1.*too few"):
        read_csv_device(b"a,b\n1,2\n3\n", dtype=np.float64)
    with pytest.raises(RuntimeError, match="row 0, column 2.*too many"):
        read_csv_device(b"a,b\n1,2,3\n", dtype=np.float64)
    with pytest.raises(RuntimeError, match="not a number"):
        read_csv_device(b"a\n1\nzz\n", dtype={"a": np.int64})
    with pytest.raises(RuntimeError, match="missing value in an int64 column"):
        read_csv_device(b"1,2\n,3\n", dtype={"a": np.int64}, names=["a", "b"])
    with pytest.raises(RuntimeError, match="int64 range"):
        read_csv_device(b"a\n99999999999999999999\n", dtype={"a": np.int64})


def test_frame_read_csv_engines(tmp_path):
    """DeviceFrame.read_csv: the device engine for numeric files, Arrow's reader when a string column shows up."""
    import pandas as pd
    from sdc_b200.frame import DeviceFrame
    rng = np.random.default_rng(8)
    df = pd.DataFrame({"k": rng.integers(0, 50, 5000), "v": rng.standard_normal(5000), "w": rng.random(5000)})
    path = tmp_path / "num.csv"
    df.to_csv(path, index=False)
    for engine in ("auto", "device", "arrow"):
        got = DeviceFrame.read_csv(str(path), engine=engine).to_pandas()
        pd.testing.assert_frame_equal(got, pd.read_csv(path, float_precision="round_trip"), check_exact=True)
    df["s"] = ["s%d" % i for i in range(5000)]
    path2 = tmp_path / "mixed.csv"
    df.to_csv(path2, index=False)
    got = DeviceFrame.read_csv(str(path2))                       # auto: not numeric -> Arrow
    assert got.columns == ["k", "v", "w", "s"] and len(got) == 5000
    from sdc_b200.csv import NotNumericCSV
    with pytest.raises(NotNumericCSV):
        DeviceFrame.read_csv(str(path2), engine="device")
