"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): the oracle on CPU and the CUDA
path on the GPU must both reproduce the stored CSR / RHS from the stored input arrays.  One fixture is also
checked by hand arithmetic (the single-cell lifting case of python/tests/test_lifting.py:24-76)."""
import glob
import os

import numpy as np
import pytest

import problems

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def _case_for(path):
    key = os.path.basename(path)[:-4]
    for name, make in problems.ALL_CASES.items():
        if name.replace(" ", "").replace("(", "").replace(")", "").replace(",", "_") == key:
            return make()
    raise KeyError(key)


def _check_inputs(c, g):
    assert np.array_equal(c.V.mesh.x_dofmap, g["x_dofmap"]) and np.array_equal(c.V.dofmap, g["dofmap"])
    assert np.allclose(c.V.mesh.x, g["x"], rtol=0, atol=1e-15)
    for k, a in zip(("slaves", "masters", "coeffs", "owners", "offsets"), c.data):
        assert np.array_equal(np.asarray(a), g[k]), k


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden(oracle, path):
    g = np.load(path)
    c = _case_for(path)
    _check_inputs(c, g)
    m = oracle.mpc_from_arrays(c.V, c.data)
    rp, col, val = oracle.assemble_matrix(c.a, m, bcs=c.bcs)
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(col, g["col"])
    assert np.allclose(val, g["val"], rtol=1e-13, atol=1e-13 * np.abs(g["val"]).max())
    if "b" in g:
        b = oracle.assemble_vector(c.L, m)
        if c.a_lift is not None and c.bcs:
            oracle.apply_lifting(b, [c.a_lift], [c.bcs], m)
        assert np.allclose(b, g["b"], rtol=1e-13, atol=1e-13 * np.abs(g["b"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_reproduces_golden(path):
    import dolfinx_mpc_b200 as mpcx

    g = np.load(path)
    c = _case_for(path)
    _check_inputs(c, g)
    mpc = mpcx.MultiPointConstraint(c.V)
    mpc.add_constraint(c.V, *c.data)
    mpc.finalize()
    A = mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs)
    rp, col, val = A.getValuesCSR()
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(col, g["col"])
    assert np.abs(val - g["val"]).max() <= 1e-10 * np.abs(g["val"]).max()
    if "b" in g:
        b = mpcx.assemble_vector(c.L, mpc)
        if c.a_lift is not None and c.bcs:
            mpcx.apply_lifting(b, [c.a_lift], [c.bcs], mpc)
        assert np.abs(b.array - g["b"]).max() <= 1e-10 * np.abs(g["b"]).max()


def test_single_quad_lifting_by_hand():
    """python/tests/test_lifting.py:24-76 worked by hand: Q1 Laplace on the unit square is
    A_e = 1/6 [[4,-1,-1,-2],[-1,4,-2,-1],[-1,-2,4,-1],[-2,-1,-1,4]] (vertices (0,0),(1,0),(0,1),(1,1));
    Dirichlet dofs 1 and 3 (x = 1), slave 0 -> master 2 with coefficient 1.  After bc zeroing and K^T A K:
    A[2,2] = A_e[2,2] + A_e[0,0] + A_e[0,2] + A_e[2,0] = (4 + 4 - 1 - 1)/6 = 1, slave and bc rows are unit rows;
    lifting gives b[2] -= (A_e[2,1] + A_e[2,3] + A_e[0,1] + A_e[0,3]) * 2.3 = -(-2-1-1-2)/6 * 2.3 = 2.3."""
    g = np.load([p for p in GOLD if "lifting-quad" in p][0])
    n = 4
    A = np.zeros((n, n))
    rp, col, val = g["row_ptr"], g["col"], g["val"]
    for r in range(n):
        A[r, col[rp[r]:rp[r + 1]]] = val[rp[r]:rp[r + 1]]
    assert np.allclose(A, np.eye(4), atol=1e-14)
    c = problems.case_lifting_single_quad()
    # b = K^T (f-part) + lifting; remove the source part with a second oracle-free evaluation: M_e f summed
    from dolfinx_mpc_b200 import elements as el

    tab = el.element_tables("quadrilateral", 1, 2)
    f = c.L.integrals[0].coefficients[0].array
    M = np.einsum("q,qi,qj->ij", tab.weights, tab.phi, tab.phi)  # unit square: detJ = 1
    be = M @ f
    expected = np.array([0.0, be[1], be[2] + be[0] + 2.3, be[3]])
    expected[1] -= A_e_row(1) @ np.array([0, 2.3, 0, 2.3])  # lifting also lands on the bc rows (set_bc overwrites later)
    expected[3] -= A_e_row(3) @ np.array([0, 2.3, 0, 2.3])
    assert np.allclose(g["b"], expected, atol=1e-13)


def A_e_row(i):
    return (np.array([[4, -1, -1, -2], [-1, 4, -2, -1], [-1, -2, 4, -1], [-2, -1, -1, 4]]) / 6.0)[i]
