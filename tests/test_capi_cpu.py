"""CPU: the C-ABI shared library builds, loads and exports every symbol include/csn_b200.h declares,
and argument validation works without a GPU (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from csn_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    header = (ROOT / "include" / "csn_b200.h").read_text()
    names = sorted(set(re.findall(r"\b(csn_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 18, names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/csn_b200.h but not exported"


def test_ctypes_signatures_cover_the_header(lib):
    from csn_b200 import _lib
    header = (ROOT / "include" / "csn_b200.h").read_text()
    names = set(re.findall(r"\bint (csn_[a-z0-9_]+)\s*\(", header))
    bound = set(_lib._EXTRA_SIGNATURES) | {"csn_gemm", "csn_abi_version", "csn_csa_head_grid"}
    assert names <= bound, names - bound


def test_abi_version_and_error_string(lib):
    assert lib.csn_abi_version() >= 1
    # argument validation happens before any CUDA call: a NULL operand is rejected with a message
    rc = lib.csn_normalize_rows(None, None, 4, 256, C.c_float(1e-12), 1, None)
    assert rc != 0
    assert b"null" in lib.csn_last_error()
    rc = lib.csn_topk_rows(C.c_void_p(16), 8, 1, 4, 65, C.c_void_p(16), C.c_void_p(16), None)
    assert rc != 0 and b"k=65" in lib.csn_last_error()
    rc = lib.csn_topk_rows(C.c_void_p(16), 8, 1, 4, 9, C.c_void_p(16), C.c_void_p(16), None)   # k > number of columns
    assert rc != 0 and b"exceeds" in lib.csn_last_error()


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    import torch
    from csn_b200 import midfc, _lib
    m = midfc.get_model("ssa", 4, 1)
    x = torch.zeros(1, 256, 10000, 1)
    with pytest.raises(_lib.CsnError):
        m(x, "test")


def test_product_code_does_not_import_the_oracle():
    for f in (ROOT / "csn_b200").glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
