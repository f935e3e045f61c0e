"""The C-ABI library builds, loads and exports every symbol ``include/clr_b200.h`` declares.
No compute calls here (no GPU in the dev container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "clr_b200.h")


@pytest.fixture(scope="module")
def lib():
    from uda_clr_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "libclr_b200.so does not export %s" % n


def test_binding_table_matches_header():
    from uda_clr_b200 import _lib
    assert _lib.exported_symbols() == declared_symbols()


def test_version_and_status_strings(lib):
    assert lib.clr_version() == 100
    assert lib.clr_status_string(0) == b"ok"
    assert b"workspace" in lib.clr_status_string(-3)
    assert lib.clr_status_string(-1000 - 700)  # CUDA error range resolves to a string


def test_workspace_queries_need_no_gpu(lib):
    assert lib.clr_pool_ws_bytes(8, 256, 128 * 128, 2) > 0
    assert lib.clr_pool_ws_bytes(0, 256, 128 * 128, 2) == 0
    assert lib.clr_pool_ws_bytes(8, 256, 128 * 128, 9) == 0


def test_argument_validation_without_gpu(lib):
    # bad arguments are rejected before any CUDA call
    assert lib.clr_pool_fwd(None, None, 0, 1, 1, 1, 2, None, 0, None, None) == -1
    assert lib.clr_proto_finalize(None, 4, 8, None, None) == -1


def test_ops_refuse_cpu_tensors():
    import torch
    import uda_clr_b200 as clr
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clr.gen_prototype(torch.zeros(1, 2, 4, 4), torch.zeros(1, 3, 4, 4))


def test_library_is_plain_c_abi():
    """No torch / Python symbols in the dynamic dependencies of the shared object."""
    from uda_clr_b200 import _lib
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out
