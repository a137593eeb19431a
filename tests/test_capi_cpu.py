"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares,
and rejects bad arguments before touching a device.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as ge
    ge.build()
    import eco_dqn_b200
    return eco_dqn_b200.lib()


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(eco_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol(L):
    import eco_dqn_b200._lib as _lib
    raw = C.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(raw, name), "libecodqn_b200.so does not export %s" % name
    assert set(_lib.EXPORTED) <= set(decl)
    assert L.eco_abi_version() == 1


def test_struct_layouts_match_header():
    import eco_dqn_b200._lib as _lib
    assert C.sizeof(_lib.Episode) == 96
    assert C.sizeof(_lib.Graphs) == 16 + 8 * 8
    assert C.sizeof(_lib.Env) == 32 + 8 + 13 * 8
    assert C.sizeof(_lib.Mpnn) == 13 * 8


def test_workspace_sizes_are_host_side_and_monotone(L):
    a = L.eco_env_workspace_bytes(64, 200, 400)
    b = L.eco_env_workspace_bytes(128, 200, 400)
    assert 0 < a < b
    assert L.eco_graphs_workspace_bytes(1, 200) >= 208 * 208
    assert L.eco_graphs_workspace_bytes(0, 200) == 0 and L.eco_graphs_workspace_bytes(1, 4096) == 0
    assert L.eco_env_workspace_bytes(1, 20, 0) == 0
    assert L.eco_mpnn_scratch_bytes(4096, 200, 1) > 0


def test_bad_arguments_are_rejected_before_any_launch(L):
    import eco_dqn_b200._lib as _lib
    g = _lib.Graphs()
    assert L.eco_graphs_bind(C.byref(g), None, 1, 20) == _lib.ECO_ERR_INVALID
    assert b"null" in L.eco_last_error()
    assert L.eco_graphs_bind(C.byref(g), C.c_void_p(256), 1, 5000) == _lib.ECO_ERR_INVALID
    assert L.eco_graphs_bind(C.byref(g), C.c_void_p(257), 1, 20) == _lib.ECO_ERR_INVALID     # alignment
    assert L.eco_graphs_bind(C.byref(g), C.c_void_p(4096), 3, 20) == 0 and g.NP == 32 and g.G == 3
    e = _lib.Env()
    assert L.eco_env_bind(C.byref(e), C.c_void_p(4096), 4, 20, 70000, -1.0) == _lib.ECO_ERR_INVALID
    assert L.eco_env_bind(C.byref(e), C.c_void_p(4096), 4, 20, 40, 0.05) == 0
    assert (e.NP, e.NW, e.HCAP, e.use_basin) == (32, 1, 128, 1)
    e2 = _lib.Env()
    assert L.eco_env_bind(C.byref(e2), C.c_void_p(4096), 4, 40, 80, -1.0) == 0 and e2.use_basin == 0
    # mismatched graph set / env sizes
    assert L.eco_env_step(C.byref(g), C.byref(e2), 0, C.c_void_p(8), None, None, None, None, None, None) == _lib.ECO_ERR_INVALID
    assert L.eco_env_step(C.byref(g), C.byref(e), 1, None, None, None, None, None, None, None) == _lib.ECO_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(_lib.ECO_ERR_INVALID)
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ECO_ERR_UNSUPPORTED)


def test_host_tables_follow_reference_fp64_arithmetic():
    import eco_dqn_b200.engine as engine
    from oracle.spin_env import time_since_flip_table, immanency_table
    for T in (40, 400, 1000):
        assert np.array_equal(engine.time_since_flip_table(T), time_since_flip_table(T))
        assert np.array_equal(engine.immanency_table(T), immanency_table(T))
    assert engine.immanency_table(40)[1] == 0.025000000000000022      # SURVEY.md appendix A.2 [probe]


def test_graph_validation_host_side():
    import eco_dqn_b200.engine as engine
    with pytest.raises(NotImplementedError):
        engine.graphs_to_int8(np.array([[0, 0.25], [0.25, 0]]))
    with pytest.raises(ValueError):
        engine.graphs_to_int8(np.zeros((2, 3, 4)))
    out = engine.graphs_to_int8([np.array([[0., -1.], [-1., 0.]])])
    assert out.dtype == np.int8 and out.shape == (1, 2, 2)


def test_no_cuda_means_loud_failure():
    import torch
    import eco_dqn_b200.engine as engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        engine.GraphSet(np.array([[0, 1], [1, 0]])[None])


def test_small_int_division_identity(tmp_path):
    """env_step_device.cuh::small_div (observable row 1 without a table or fp64): gain / mlr by the correctly rounded
    reciprocal and two FMAs equals the reference's fp64 division followed by the fp32 cast, for every |mlr| <= 2048 and
    |gain| <= 70000 -- exhaustively, in C (oracle/small_div_check.c)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "small_div_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", os.path.join(root, "oracle", "small_div_check.c"), "-o", exe, "-lm"],
                   check=True)
    out = subprocess.run([exe, "2048", "70000"], check=True, capture_output=True, text=True, timeout=600).stdout.split()
    assert int(out[0]) == 2 * 2048 * (2 * 70000 + 1) and int(out[1]) == 0, out
