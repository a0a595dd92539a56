"""Host side of the Arrow bridge: Arrow string buffers ARE the reference's str_arr layout
(sdc/str_arr_type.py:33-101), so `StringArray.from_arrow` must agree with the element-by-element boxing of
`from_pylist` -- slices, nulls, large_string, dictionary-encoded and chunked inputs included."""
import numpy as np
import pyarrow as pa
import pytest

from sdc_b200.str_arr import StringArray

VALUES = ["b", None, "", "héllo", "zz", None, "a" * 40, "tail"]


def _same(a, b):
    assert a.n == b.n
    np.testing.assert_array_equal(a.offsets, b.offsets)
    np.testing.assert_array_equal(a.chars, b.chars)
    np.testing.assert_array_equal(a.bitmap[: (a.n + 7) // 8], b.bitmap[: (b.n + 7) // 8])
    assert a.offsets.dtype == np.uint32 and a.chars.dtype == np.uint8


@pytest.mark.parametrize("typ", [pa.string(), pa.large_string()])
def test_from_arrow_matches_boxing(typ):
    arr = pa.array(VALUES, type=typ)
    _same(StringArray.from_arrow(arr), StringArray.from_pylist(VALUES))
    _same(StringArray.from_arrow(arr.slice(2, 5)), StringArray.from_pylist(VALUES[2:7]))          # non-zero offset
    _same(StringArray.from_arrow(pa.chunked_array([arr.slice(0, 3), arr.slice(3)])), StringArray.from_pylist(VALUES))
    _same(StringArray.from_arrow(arr.dictionary_encode()), StringArray.from_pylist(VALUES))
    _same(StringArray.from_arrow(pa.array([], type=typ)), StringArray.from_pylist([]))
    no_nulls = ["x", "yy", ""]
    _same(StringArray.from_arrow(pa.array(no_nulls, type=typ)), StringArray.from_pylist(no_nulls))


def test_bytes_under_null_slots_are_dropped():
    # Arrow allows data bytes under a null slot; the reference stores NA strings empty (sdc_str_arr.hpp:389-393)
    offsets = pa.py_buffer(np.array([0, 2, 5, 6], dtype=np.int32).tobytes())
    data = pa.py_buffer(b"abXYZc")
    validity = pa.py_buffer(np.packbits([1, 0, 1], bitorder="little").tobytes())
    arr = pa.StringArray.from_buffers(3, offsets, data, validity, 1)
    _same(StringArray.from_arrow(arr), StringArray.from_pylist(["ab", None, "c"]))


def test_round_trip_to_arrow():
    back = StringArray.from_pylist(VALUES).to_arrow()
    assert back.to_pylist() == VALUES
    with pytest.raises(TypeError):
        StringArray.from_arrow(pa.array([1, 2, 3]))
