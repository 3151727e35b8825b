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


def header_prototypes():
    """{name: [parameter type strings]} parsed from the header's declarations."""
    text = open(os.path.join(ROOT, "include", "pixie_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(ln for ln in text.splitlines() if not ln.lstrip().startswith("#"))
    out = {}
    for m in re.finditer(r"\b(pixie_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if params in ([""], ["void"]) else params
    return out


def test_python_binding_matches_the_header_parameter_by_parameter():
    """Every ctypes argtypes list has the header's arity and each parameter the header's kind
    (pointer / 32-bit / 64-bit / double / size_t): a drifted binding would otherwise only show on
    the GPU box."""
    L = _native.lib()
    protos = header_prototypes()
    assert sorted(protos) == header_symbols()
    c = ctypes

    def kind(decl):
        if "*" in decl:
            return {c.c_void_p, c.c_char_p}
        t = decl.rsplit(" ", 1)[0].replace("const", "").strip()
        return {"int32_t": {c.c_int32}, "int": {c.c_int, c.c_int32}, "uint32_t": {c.c_uint32},
                "int64_t": {c.c_int64}, "double": {c.c_double}, "size_t": {c.c_size_t},
                "float": {c.c_float}}[t]

    checked = 0
    for name, params in protos.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            assert not params or name in ("pixie_version", "pixie_device_count",
                                          "pixie_kernel_launches"), name
            continue
        assert len(fn.argtypes) == len(params), (name, len(fn.argtypes), params)
        for got, decl in zip(fn.argtypes, params):
            assert got in kind(decl), (name, decl, got)
        checked += 1
    assert checked >= 15


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
