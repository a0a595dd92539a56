"""The C-ABI library loads and exports every symbol include/sdc_b200.h declares (no GPU needed)."""
import os
import re

from sdc_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    """Function names declared by the header after preprocessing (macros expanded)."""
    import subprocess
    text = subprocess.run(["gcc", "-E", "-P", os.path.join(ROOT, "include", "sdc_b200.h")],
                          check=True, capture_output=True, text=True).stdout
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
    return sorted({n for n in names if not n.startswith("__")})


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_signature_table_matches_header():
    assert sorted(native.SIGNATURES) == _declared_symbols()


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.sdcb200_version()
    assert isinstance(lib.sdcb200_last_error(), bytes)


def test_fn_address_nonzero():
    assert native.fn_address("sdcb200_rolling")
