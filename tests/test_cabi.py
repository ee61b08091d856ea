"""CPU checks of the drop-in boundary: libmpcx.so loads without a GPU and exports every symbol that
include/mpcx.h declares; argument validation that needs no device works; the product package never
touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mpcx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcx_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from dolfinx_mpc_b200 import _lib

    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from dolfinx_mpc_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"libmpcx.so does not export {name}"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding and header disagree"
    assert lib.mpcx_abi_version() == 6


def test_null_arguments_are_rejected_without_a_device(lib):
    from dolfinx_mpc_b200 import _lib

    rc = lib.mpcx_assemble_matrix_f64(None, None, None, None, None, None, None, None, None, None, None)
    assert rc == _lib.ERR_ARG
    assert b"null" in lib.mpcx_last_error()
    assert lib.mpcx_backsubstitution_f64(None, None, None) == _lib.ERR_ARG
    with pytest.raises(_lib.MpcxError):
        _lib.check(lib.mpcx_build_plan(None, None, None, 0, None, None, 1, None))


def test_host_pattern_matches_oracle(lib, oracle):
    """create_sparsity_pattern (host C++, threaded) vs the oracle's restatement of cpp/utils.h:381-496:
    bit-exact row_ptr / col on every fixture."""
    import numpy as np

    import problems
    from dolfinx_mpc_b200 import MultiPointConstraint, create_sparsity_pattern

    for name, make in problems.ALL_CASES.items():
        c = make()
        mpc = MultiPointConstraint(c.V)
        mpc.add_constraint(c.V, *c.data)
        mpc.finalize()
        rp, col = create_sparsity_pattern(c.a, mpc, num_threads=3)
        m = oracle.mpc_from_arrays(c.V, c.data)
        rp_o, col_o = oracle.create_pattern(c.a, m, m)
        assert np.array_equal(rp, rp_o) and np.array_equal(col, col_o), name


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import problems
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import _lib

    c = problems.case_periodic_2d(4, 1, False)
    mpc = mpcx.MultiPointConstraint(c.V)
    mpc.add_constraint(c.V, *c.data)
    mpc.finalize()
    with pytest.raises(_lib.MpcxError):
        mpcx.assemble_matrix(c.a, mpc)
    with pytest.raises(_lib.MpcxError):
        mpcx.assemble_vector(c.L, mpc)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dolfinx_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "libmpc_oracle", "mpc_oracle.c", "orc_"):
                    assert needle not in text, f"{f} reaches into the oracle ({needle})"
