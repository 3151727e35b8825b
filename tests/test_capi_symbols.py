"""CPU test: libpixie_b200.so builds (nvcc cross-compiles sm_100a without a GPU), loads, and
exports every symbol include/pixie_b200.h declares.  No compute calls are made here."""
import ctypes
import os
import re
import subprocess

from ark_analysis_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pixie_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pixie_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _native.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = header_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pixie_b200.h but not exported"
    # the Python binding knows exactly the same set
    assert sorted(_native.SYMBOLS) == declared


def test_version_error_strings_and_workspace_size():
    L = _native.lib()
    assert L.pixie_version() >= 100
    assert L.pixie_error_string(0) == b"ok"
    assert b"workspace" in L.pixie_error_string(-2)
    small = L.pixie_workspace_bytes(0, 32, 100)
    big = L.pixie_workspace_bytes(1 << 20, 32, 100)
    assert 0 < small < big
    assert L.pixie_workspace_bytes(-1, 32, 100) == 0


def test_binary_is_blackwell_native():
    """SASS must hold the tcgen05 / TMA / TMEM instructions (UTC*MMA, UTMALDG, LDTM)."""
    sass = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True,
                          text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass
