"""CPU: csrc/num_parse.cuh (the number conversion the CSV kernels compile) built for the host with g++ and checked
field by field against Python's correctly rounded float() / int() -- bit-exact -- and against Arrow's grammar."""
import math
import os
import random
import shutil
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("np") / "num_parse_harness")
    subprocess.run(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "num_parse_harness.cpp")], check=True)
    return exe


def _run(exe, mode, fields):
    out = subprocess.run([exe, mode], input="\n".join(fields) + "\n", capture_output=True, text=True, check=True).stdout
    return [ln.split() for ln in out.splitlines()]


def _bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def test_pow5_table_is_what_the_generator_writes():
    gen = subprocess.run(["python", os.path.join(ROOT, "tools", "gen_pow5_table.py")], capture_output=True, text=True, check=True).stdout
    assert gen == open(os.path.join(ROOT, "sdc_b200", "csrc", "pow5_table.inc")).read()


def test_float_fields_bit_exact(harness):
    rng = random.Random(7)
    fields = ["1.5", "+1.5", "-1.5", " 1.5", "1.5\t", "1e5", "1E+5", ".5", "5.", "-.5e-3", "1e400", "-1e400", "1e-400",
              "1.7976931348623157e308", "1.7976931348623159e308", "4.9e-324", "2.4703282292062327e-324",
              "2.4703282292062328e-324", "2.2250738585072014e-308", "2.2250738585072011e-308", "00012", "0e0", "-0.0",
              "9007199254740993", "9007199254740992.5", "9007199254740991.5", "1e23", "8.41e21", "1e22", "1e-22",
              "7.2057594037927933e16", "0.000000000000000000000000000000000000001", "123456789012345678", "0.0e-999999999",
              "18446744073709551615", "18446744073709551616", "9999999999999999999", "99999999999999999999"]
    for _ in range(200_000):
        k = rng.random()
        if k < 0.35:
            x = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
            if math.isnan(x) or math.isinf(x):
                continue
            fields.append(repr(x))
        elif k < 0.55:
            fields.append("%.*f" % (rng.randint(0, 12), rng.uniform(-1e6, 1e6)))
        elif k < 0.8:
            nd = rng.randint(1, 19)
            frac = "" if rng.random() < 0.5 else "." + str(rng.randint(0, 999))
            fields.append("%d%se%d" % (rng.randint(0, 10 ** nd - 1), frac, rng.randint(-330, 300)))
        elif k < 0.9:
            # exactly halfway between two doubles, at most 19 digits: round to even
            m = rng.getrandbits(53) | (1 << 52)
            e = rng.randint(0, 10)
            v = (2 * m + 1) << e
            if len(str(v)) <= 19:
                fields.append(str(v))
                fields.append("%d.0" % v)
        else:
            fields.append("%.25e" % struct.unpack("<d", struct.pack("<Q", rng.getrandbits(62)))[0])
    got = _run(harness, "f", fields)
    assert len(got) == len(fields)
    for s, (st, b) in zip(fields, got):
        assert st == "0", (s, st)
        assert int(b, 16) == _bits(float(s)), (s, b, "%016x" % _bits(float(s)))


def test_float_fields_longer_than_19_digits(harness):
    """More than 19 significant digits: the 19-digit truncation decides unless its bracket straddles a rounding boundary;
    then the whole digit string is compared with the midpoint of the two candidate doubles in exact integer arithmetic
    (num_parse.cuh round_long_decimal).  Exact ties, a hair above / below a tie, subnormals, the overflow boundary,
    digit strings beyond the 800 digits that take part."""
    from fractions import Fraction
    rng = random.Random(11)
    fields = []
    for _ in range(20_000):
        m = rng.getrandbits(53) | (1 << 52)
        e = rng.randint(-1100, 960)
        if rng.random() < 0.15:
            m, e = rng.getrandbits(52), -1074                      # subnormal
        v = Fraction(2 * m + 1, 2) * (Fraction(2) ** e)            # halfway between two adjacent doubles
        k2 = v.denominator.bit_length() - 1
        digits = str(v.numerator * 5 ** k2)
        fields.append(digits + ("e-%d" % k2 if k2 else ""))
        if rng.random() < 0.5:
            fields.append(digits + "0000001" + "e-%d" % (k2 + 7))
        else:
            fields.append(str(v.numerator * 5 ** k2 * 10 ** 7 - 1) + "e-%d" % (k2 + 7))
    for _ in range(40_000):
        x = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        if math.isnan(x) or math.isinf(x):
            continue
        fields.append("%.30e" % x)
    big = ("1.797693134862315807937289714053034150799341327100378269361737789804449682927647509466490179775872070963302864166"
           "9288791094655554785194040263065748867150582068190890200070838367627385484581771153176447573027006985557136695962"
           "284291481986083493647529271907416844436551070434271155969950809304288017790417449779")
    fields += ["0." + "0" * 400 + "1" + "9" * 900, "1" + "0" * 300 + "." + "0" * 20 + "1", "9" * 1000 + "e-700",
               "0.5" + "0" * 900 + "1e-323", "2.4703282292062327208051355972476e-324", "2.4703282292062327208051355972479e-324",
               big + "e308", big[:-1] + "80e308", "12345678901234567890123", "0.1000000000000000055511151231257827"]
    got = _run(harness, "f", fields)
    assert len(got) == len(fields)
    for s, (st, b) in zip(fields, got):
        assert st == "0", (s[:80], st)
        assert int(b, 16) == _bits(float(s)), (s[:80], b, "%016x" % _bits(float(s)))


def test_float_specials_and_rejects(harness):
    nan = "7ff8000000000000"
    cases = {"inf": "7ff0000000000000", "-inf": "fff0000000000000", "+Infinity": "7ff0000000000000", "INF": "7ff0000000000000",
             "NAN": nan, "-NAN": nan, "+nan": nan}
    got = _run(harness, "f", list(cases))
    for (s, want), (st, b) in zip(cases.items(), got):
        assert st == "0" and b == want, (s, st, b)
    # Arrow's null spellings -> NaN with the "null" status
    nulls = ["", "NA", "N/A", "n/a", "NaN", "nan", "-NaN", "-nan", "NULL", "null", "#N/A", "#NA", "#N/A N/A", "1.#IND", "1.#QNAN",
             "-1.#IND", "-1.#QNAN"]
    got = _run(harness, "f", nulls)
    assert [g[0] for g in got] == ["1"] * len(nulls) and all(g[1] == nan for g in got)
    bad = ["0x10", "1_000", "1e", "1e+", "e5", "-", "+", "1.2.3", ".", "1 2", "abc", "1,5", "--1", "1e5.5", "Na"]
    got = _run(harness, "f", bad)
    assert [g[0] for g in got] == ["2"] * len(bad), list(zip(bad, got))


def test_int_fields(harness):
    ok = {"12": 12, "-12": -12, "012": 12, " 12": 12, "7\t": 7, "0": 0, "-0": 0, "9223372036854775807": 2**63 - 1,
          "-9223372036854775808": -2**63}
    got = _run(harness, "i", list(ok))
    for (s, want), (st, v) in zip(ok.items(), got):
        assert st == "0" and int(v) == want, (s, st, v)
    rng = random.Random(3)
    vals = [rng.randint(-2**63, 2**63 - 1) for _ in range(50_000)]
    got = _run(harness, "i", [str(v) for v in vals])
    assert all(st == "0" and int(v) == w for (st, v), w in zip(got, vals))
    status = {"+12": "2", "1.0": "2", "1e3": "2", "-": "2", "": "1", "NA": "1", "9223372036854775808": "3",
              "-9223372036854775809": "3", "18446744073709551616": "3", "99999999999999999999": "3"}
    got = _run(harness, "i", list(status))
    assert [g[0] for g in got] == list(status.values()), list(zip(status, got))
