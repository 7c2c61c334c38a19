"""CPU-only: libccx.so builds for sm_100a, loads without a GPU and exports every symbol that
include/ccx.h declares; the Python binding table matches the header; the product has no CPU path."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from chinesecheckersagent_b200 import build
    return build.build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ccx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccx_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = header_symbols()
    for must in ("ccx_create", "ccx_movegen", "ccx_apply", "ccx_step_random", "ccx_encode", "ccx_play_greedy"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    L = ctypes.CDLL(lib_path)
    for s in header_symbols():
        assert hasattr(L, s), "libccx.so lacks %s declared in include/ccx.h" % s
    assert L.ccx_abi_version() == 2


def test_python_binding_table_matches_header():
    from chinesecheckersagent_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    _lib.load()


def test_sass_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    from chinesecheckersagent_b200 import _lib
    from chinesecheckersagent_b200.engine import Engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.CcxError):
        Engine()


def test_argument_errors_are_reported_not_crashed(lib_path):
    from chinesecheckersagent_b200 import _lib
    L = _lib.load()
    assert L.ccx_movegen(None, 1, None, None) == -1
    assert L.ccx_create(0, None) == -1
    assert L.ccx_strerror(-1) == b"invalid argument"
    assert L.ccx_destroy(None) == 0


def test_product_never_imports_oracle():
    """The product must not import, link or execute anything under oracle/ (nor the reference shim)."""
    pkg = os.path.join(ROOT, "chinesecheckersagent_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|import\s+refshim|ccx_oracle|libccx_oracle|orc_[a-z]+\(|oracle/|hostcheck\.")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), "%s references the oracle" % f


def test_sass_uses_the_blackwell_tensor_path(lib_path):
    """SASS evidence (B200_PROFILING.md): tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, bulk async copies -> UBLKCP,
    and no legacy HMMA (mma.sync / wmma) anywhere in the library."""
    out = subprocess.run(["cuobjdump", "-sass", lib_path], stdout=subprocess.PIPE, text=True).stdout
    counts = {m: len(re.findall(r"\b%s\b" % m, out)) for m in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR")}
    assert all(v > 0 for v in counts.values()), counts
    assert not re.search(r"\bHMMA\b", out)
