"""CPU: the C-ABI library builds, loads without a GPU, and exports exactly what include/kssd_b200.h declares.
No compute call is made here."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built():
    from public_kssd_b200 import capi
    capi.build_library()
    return capi


def _header_functions():
    txt = (ROOT / "include" / "kssd_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kssd_[a-z0-9_]+)\s*\(", txt)))


def test_header_matches_binding_list(built):
    assert _header_functions() == sorted(built.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    out = subprocess.run(["nm", "-D", "--defined-only", str(built.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (kssd_[a-z0-9_]+)", out))
    missing = [s for s in _header_functions() if s not in exported]
    assert not missing, missing
    lib = built.lib()
    for s in built.SYMBOLS:
        assert hasattr(lib, s)
    assert b"sm_100a" in lib.kssd_version()


def test_sass_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", str(built.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device(built):
    """On a box without a GPU the product path must fail loudly, not compute on the CPU."""
    import numpy as np
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    from public_kssd_b200 import kssd
    with pytest.raises(kssd.KssdError) as e:
        kssd.Context(8, 5, 2, np.arange(1 << 20, dtype=np.int32))
    assert e.value.code == -2


def test_product_never_imports_oracle():
    for p in (ROOT / "public_kssd_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h") and p.is_file():
            assert "oracle" not in p.read_text().replace("oracle restatement", ""), p
