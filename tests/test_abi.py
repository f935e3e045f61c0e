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
    assert lib.clr_version() == 200
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


def test_transnorm_argument_validation_without_gpu(lib):
    # clr_tn_*: workspace query and argument checks return before any CUDA call
    assert lib.clr_tn_ws_bytes(305) > 0 and lib.clr_tn_ws_bytes(0) == 0
    assert lib.clr_tn_fwd(None, 4, 8, 16, None, None, None, None, None, None, 0.1, 1e-5, None, 0, None, None, None) == -1
    assert lib.clr_tn_bwd(None, None, 4, 8, 16, None, None, 0, None, 0, None, None, None, None) == -1
    assert lib.clr_tn_eval(None, 4, 8, 16, None, None, None, None, None, None, 1e-5, None, 0, None, None, None) == -1


def test_transnorm_module_refuses_cpu_tensors():
    import torch
    from uda_clr_b200 import TransNorm2d
    m = TransNorm2d(4)
    assert list(m.state_dict()) == ["weight", "bias", "running_mean_source", "running_var_source", "running_mean_target",
                                    "running_var_target", "num_batches_tracked"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 4, 8, 8))


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


def test_new_entry_points_validate_arguments_without_gpu(lib):
    """Argument checks of the one-pass / glue / exchange entry points run before any CUDA call."""
    import ctypes
    assert lib.clr_mc_retrify(None, None, 8, 1, 2, 8, 8, 32, 32, 0.75, 0.04, None, None, None, None, None) == -1
    assert lib.clr_seg_loss_ws_bytes() > 0
    assert lib.clr_seg_loss_fwd(None, None, 0, None, None, 0, None, 0, None, None) == -1
    assert lib.clr_seg_loss_bwd(None, None, 0, None, None, 0, None, 1.0, None, None, None) == -1
    assert lib.clr_entropy_fwd(None, 0, 1e-7, None, None) == -1
    assert lib.clr_seg_counts(None, None, 1, 2, 16, 0.75, None, None) == -1
    assert lib.clr_pool_rows_fwd_ps(None, None, 1, 1, 1, 1, None, 0, None, None) == -3      # workspace check comes first
    assert lib.clr_bmm_finalize(None, 1, 1, 1, 1.0, None, None) == -1
    assert lib.clr_peer_alloc(0, None) == -1 and lib.clr_peer_export(None, None) == -1
    assert lib.clr_peer_open(None, None) == -1 and lib.clr_peer_close(None) == -1 and lib.clr_peer_free(None) == -1
    assert lib.clr_trace_slots() >= 16 and lib.clr_trace_name(0) == b"mc_stats"


def test_exchange_buffer_size_formula(lib):
    """Receive buffer of the in-kernel exchange: 64-bit words, 2 parities x world sources x (packed1 + packed2 + tail)."""
    K, C = 2, 256
    n1, n2 = 2 * 2 * K * (C + 1), K * (C + 1) + 4
    for world in (1, 2, 8):
        assert lib.clr_step_xchg_bytes(world, K, C) == 8 * 2 * world * (n1 + n2)
    assert lib.clr_step_xchg_bytes(9, K, C) == 0 and lib.clr_step_xchg_bytes(2, 9, C) == 0


def test_step_args_layout_matches_header():
    """The ctypes mirror of clr_step_args must keep the header's field order (the exchange fields were appended)."""
    from uda_clr_b200 import _lib
    src = open(HEADER).read()
    start = src.index("typedef struct clr_step_args {") + len("typedef struct clr_step_args {")
    body = src[start:src.index("} clr_step_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            m = re.search(r"(\w+)\s*(\[\w+\])?\s*$", part.strip())
            if m:
                names.append(m.group(1))
    assert names == [f[0] for f in _lib.StepArgs._fields_]


def test_header_is_plain_c99_and_c_host_links(lib, tmp_path):
    """The boundary is a C ABI: the header must compile as C (not only C++), and a C host must link against the
    library with nothing but cudart (examples/clr_host.c; it is RUN in tests/test_gpu_parity.py)."""
    import shutil
    import subprocess
    from uda_clr_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", "-Wall", "-Wpedantic", HEADER], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA toolkit headers not found")
    exe = str(tmp_path / "clr_host")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
                        os.path.join(ROOT, "examples", "clr_host.c"), "-o", exe, "-L", os.path.dirname(_lib.LIB_PATH),
                        "-lclr_b200", "-L", cuda_lib, "-lcudart", "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
