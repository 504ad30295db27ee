"""CPU: the C-ABI library loads and exports every symbol include/agenda_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from agenda_b200 import build
    return build.build()


def _declared():
    text = open(os.path.join(ROOT, "include", "agenda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agenda_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/agenda_b200.h but not exported"


def test_binding_covers_header(built):
    from agenda_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()
    lib = _lib.load()
    assert lib.agenda_version() == 1000
    assert lib.agenda_last_error() == b""


def test_bad_arguments_fail_without_gpu(built):
    """Argument validation happens before any CUDA call, so it is checkable on the CPU box."""
    from agenda_b200 import _lib
    lib = _lib.load()
    rc = lib.agenda_heat_normalize_u8(None, None, 1, 16, None)
    assert rc == -1 and b"null" in lib.agenda_last_error()
    rc = lib.agenda_ccl_bbox(ctypes.c_void_p(16), 0.5, None, None, None, 0, 1, -3, 4, None)
    assert rc == -2
    rc = lib.agenda_attn_self_fwd(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16),
                                  0, 1, 1, 128, 40, 0.1, None)
    assert rc == -3 and b"bf16" in lib.agenda_last_error()
    with pytest.raises(_lib.AgendaError):
        _lib.call("agenda_heat_finalize", 16, 16, 4, 0, None)


def test_product_has_no_cpu_fallback():
    import torch
    from agenda_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.heat_normalize_u8(torch.zeros(2, 4, 4))
    # nothing under agenda_b200/ imports the oracle
    pkg = os.path.join(ROOT, "agenda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f


def test_binding_arity_matches_header():
    """Every ctypes signature in agenda_b200/_lib.py has as many arguments as the prototype in include/agenda_b200.h
    (a mismatch would not fail at load time: ctypes would silently pass garbage)."""
    from agenda_b200 import _lib
    text = open(os.path.join(ROOT, "include", "agenda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {m.group(1): m.group(2) for m in re.finditer(r"\b(agenda_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)}
    for name, argtypes in _lib._SIGNATURES.items():
        assert name in protos, name
        params = [p for p in protos[name].split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
