"""The C-ABI library loads without a GPU and exports every symbol include/tstereo.h declares,
with the argument counts the ctypes binding (temporalstereo_b200/_lib.py) assumes."""
import ctypes
import os
import re

import pytest

from temporalstereo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tstereo.h")


def _declarations():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(tstereo_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        decls[name] = (ret, n, args)
    return decls


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_declares_what_python_binds():
    decls = _declarations()
    assert len(decls) >= 20
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, (ret, n, args) in decls.items():
        assert n == len(_lib.SIGNATURES[name][1]), f"{name}: header has {n} args"
        # every pointer in the header is a plain C pointer (no torch / C++ types cross the ABI)
        assert "Tensor" not in args and "std::" not in args and "&" not in args


def test_argument_kinds_match_header():
    kinds = {"P": ctypes.c_void_p, "I": ctypes.c_int, "F": ctypes.c_float, "L": ctypes.c_longlong}
    for name, (ret, n, args) in _declarations().items():
        if n == 0:
            continue
        got = _lib.SIGNATURES[name][1]
        for i, a in enumerate(args.split(",")):
            a = a.strip()
            if "*" in a:
                want = kinds["P"]
            elif a.startswith("long long"):
                want = kinds["L"]
            elif a.startswith("float"):
                want = kinds["F"]
            else:
                assert a.startswith("int"), (name, a)
                want = kinds["I"]
            assert got[i] is want, f"{name} arg {i} ({a})"


def test_library_exports_every_symbol(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declarations():
        assert hasattr(raw, name), name
    assert lib.tstereo_version() == 200
    assert lib.tstereo_last_error() is not None


def test_scratch_size_is_pure_host_arithmetic(lib):
    # B*G*D*((H/2)*(W/2) + (H/4)*(W/4))
    assert lib.tstereo_block_cost_scratch_floats(2, 16, 10, 14, 5) == 2 * 2 * 5 * (5 * 7 + 2 * 3)


def test_argument_validation_needs_no_gpu(lib):
    # null pointers are rejected before any CUDA call, with a message
    rc = lib.tstereo_block_cost_shift(None, None, None, None, 1, 8, 8, 8, 4, None)
    assert rc == -1
    assert b"null" in lib.tstereo_last_error()
    with pytest.raises(_lib.TStereoError):
        _lib.call("tstereo_predict_disp", None, None, None, None, None, None, 1, 4, 8, 8, None)


def test_reference_splat_library_loads():
    """oracle/_ref/libsoftsplat_ref.so (the reference's own splat kernel, built by oracle/build_softsplat_ref.py in the
    build container) loads without a GPU and exports its launcher."""
    path = os.path.join(ROOT, "oracle", "_ref", "libsoftsplat_ref.so")
    if not os.path.exists(path):
        if not os.path.isdir("/root/reference"):
            pytest.skip("no reference tree and no prebuilt oracle/_ref")
        import subprocess, sys
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "build_softsplat_ref.py")], check=True)
    assert hasattr(ctypes.CDLL(path), "softsplat_ref_launch")
