"""CPU tests: the oracle against known answers and against the reference's own test identities.

The reference cannot be imported here, so its fixtures are re-created (tests/problems.py) and checked with
the method of python/src/dolfinx_mpc/utils/test.py (K^H A K, K^H b) -- the independent second oracle.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg

import problems
from dolfinx_mpc_b200 import elements as el, fem, generators as gen
from dolfinx_mpc_b200.fem import Kernel


def _csr(t, n):
    rp, col, val = t
    return sp.csr_matrix((val, col, rp), shape=(n, n))


def test_p1_reference_triangle_stiffness(oracle):
    tab = el.element_tables("triangle", 1, 0)
    X = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=float)
    A = oracle.tabulate(Kernel.LAPLACE, tab, 1, X)
    assert np.allclose(A, [[1, -0.5, -0.5], [-0.5, 0.5, 0], [-0.5, 0, 0.5]], atol=1e-15)
    tabm = el.element_tables("triangle", 1, 2)
    M = oracle.tabulate(Kernel.MASS, tabm, 1, X)
    assert np.allclose(M, (np.ones((3, 3)) + np.eye(3)) / 24, atol=1e-15)


def test_q1_unit_square_stiffness(oracle):
    """The single-cell matrix behind python/tests/test_lifting.py: Q1 Laplace on the unit square."""
    tab = el.element_tables("quadrilateral", 1, 2)
    X = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=float)
    A = oracle.tabulate(Kernel.LAPLACE, tab, 1, X)
    ref = np.array([[4, -1, -1, -2], [-1, 4, -2, -1], [-1, -2, 4, -1], [-2, -1, -1, 4]]) / 6
    assert np.allclose(A, ref, atol=1e-14)


@pytest.mark.parametrize("cell,degree", [("triangle", 2), ("tetrahedron", 1), ("tetrahedron", 2)])
def test_element_tensor_properties(oracle, cell, degree):
    rng = np.random.default_rng(3)
    tdim = el.CELL_TDIM[cell]
    X = np.zeros((tdim + 1, 3))
    X[:, :tdim] = np.vstack([np.zeros(tdim), np.eye(tdim)]) + 0.2 * rng.random((tdim + 1, tdim))
    J = (X[1:, :tdim] - X[0, :tdim]).T
    vol = abs(np.linalg.det(J)) / {2: 2, 3: 6}[tdim]
    A = oracle.tabulate(Kernel.LAPLACE, el.element_tables(cell, degree, 2 * degree - 2), 1, X, c=[2.0])
    assert np.allclose(A, A.T) and np.allclose(A.sum(axis=1), 0, atol=1e-13)
    M = oracle.tabulate(Kernel.MASS, el.element_tables(cell, degree, 2 * degree), 1, X)
    assert abs(M.sum() - vol) < 1e-14
    # elasticity: rigid body modes are in the kernel
    tabe = el.element_tables(cell, degree, 2 * degree - 2)
    E = oracle.tabulate(Kernel.ELASTICITY, tabe, tdim, X, c=[1.5, 0.7])
    assert np.allclose(E, E.T)
    nodes_ref = np.vstack([np.zeros(tdim), np.eye(tdim)])
    if degree == 2:
        edges = {2: el._TRI_EDGES, 3: el._TET_EDGES}[tdim]
        nodes_ref = np.vstack([nodes_ref] + [(nodes_ref[a] + nodes_ref[b]) / 2 for a, b in edges])
    P = nodes_ref @ J.T + X[0, :tdim]
    for k in range(tdim):
        t = np.zeros((len(P), tdim)); t[:, k] = 1
        assert np.allclose(E @ t.reshape(-1), 0, atol=1e-12)
    rot = np.stack([-P[:, 1], P[:, 0]] + ([np.zeros(len(P))] if tdim == 3 else []), axis=1)
    assert np.allclose(E @ rot.reshape(-1), 0, atol=1e-11)
    # source vector = M f
    f = rng.random(M.shape[0])
    b = oracle.tabulate(Kernel.SOURCE, el.element_tables(cell, degree, 2 * degree), 1, X, w=f)
    assert np.allclose(b, M @ f, atol=1e-14)


def test_k_matrix_docstring_example(oracle):
    """utils/test.py:72-85: dim 3, u_1 = alpha u_0 + beta u_2 -> K = [[1,0],[alpha,beta],[0,1]]."""
    K = oracle.transformation_matrix(3, [1], [0, 2], [0.3, 0.9], [0, 2]).toarray()
    assert np.allclose(K, [[1, 0], [0.3, 0.9], [0, 1]])


def test_mpc_build_matches_product_finalize(oracle):
    """Integer maps bit-exact: oracle restatement of the constructor vs MultiPointConstraint.finalize."""
    from dolfinx_mpc_b200 import MultiPointConstraint

    for name in ("periodic3d-P2-bs1-3-ax2-bc1", "contact3d", "slip3d-P2-2", "tie2d-bs2", "empty-mpc"):
        c = problems.ALL_CASES[name]()
        m = oracle.mpc_from_arrays(c.V, c.data)
        mpc = MultiPointConstraint(c.V)
        mpc.add_constraint(c.V, *c.data)
        mpc.finalize()
        assert np.array_equal(mpc.is_slave, m.is_slave)
        assert np.array_equal(mpc.slaves, m.slaves) and mpc.num_local_slaves == m.num_local_slaves
        assert np.array_equal(mpc.masters.offsets, m.offsets) and np.array_equal(mpc.masters.array, m.masters)
        assert np.array_equal(mpc.coefficients()[0], m.coeffs)
        assert np.array_equal(mpc.cell_to_slaves.offsets, m.c2s_offsets)
        assert np.array_equal(mpc.cell_to_slaves.array, m.c2s)


@pytest.mark.parametrize("name", list(problems.ALL_CASES))
def test_oracle_satisfies_reference_identities(oracle, name):
    c = problems.ALL_CASES[name]()
    n = c.V.num_dofs
    m = oracle.mpc_from_arrays(c.V, c.data)
    e = oracle.OracleMPC.empty(c.V)
    A_mpc = _csr(oracle.assemble_matrix(c.a, m, bcs=c.bcs), n)
    A_org = _csr(oracle.assemble_matrix(c.a, e, bcs=c.bcs), n)
    slaves, masters, coeffs, _, offsets = c.data
    K = oracle.transformation_matrix(n, slaves, masters, coeffs, offsets)
    oracle.compare_mpc_lhs(A_org, A_mpc, K, slaves)
    # slave rows / columns hold only the diagonal (cpp/assemble_matrix.cpp:165-178,711-724)
    if len(slaves):
        S = A_mpc[np.asarray(slaves)]
        assert np.allclose(S.diagonal(k=0) if False else A_mpc.diagonal()[slaves], 1.0)
        assert abs(S).sum() == pytest.approx(len(slaves))
    if c.L is not None:
        b = oracle.assemble_vector(c.L, m)
        b0 = oracle.assemble_vector(c.L, e)
        if c.a_lift is not None and c.bcs:
            oracle.apply_lifting(b, [c.a_lift], [c.bcs], m)
            oracle.apply_lifting(b0, [c.a_lift], [c.bcs], e)
        oracle.compare_mpc_rhs(b0, b, K, slaves)


def test_lifting_solve_matches_reduced_system(oracle):
    """python/tests/test_lifting.py:78-119: solve the MPC system, back-substitute, compare with the reduced
    system K^T A K d = K^T b solved by scipy."""
    c = problems.case_periodic_2d(8, 1, True)
    n = c.V.num_dofs
    m = oracle.mpc_from_arrays(c.V, c.data)
    e = oracle.OracleMPC.empty(c.V)
    slaves, masters, coeffs, _, offsets = c.data
    A = _csr(oracle.assemble_matrix(c.a, m, bcs=c.bcs), n)
    b = oracle.assemble_vector(c.L, m)
    oracle.apply_lifting(b, [c.a], [c.bcs], m)
    vals = np.zeros(n); c.bcs[0].set(vals)
    b[c.bcs[0].dofs] = vals[c.bcs[0].dofs]
    u = scipy.sparse.linalg.spsolve(A.tocsc(), b)
    oracle.backsubstitution(m, u)
    A0 = _csr(oracle.assemble_matrix(c.a, e, bcs=c.bcs), n)
    b0 = oracle.assemble_vector(c.L, e)
    oracle.apply_lifting(b0, [c.a], [c.bcs], e)
    b0[c.bcs[0].dofs] = vals[c.bcs[0].dofs]
    K = oracle.transformation_matrix(n, slaves, masters, coeffs, offsets)
    d = scipy.sparse.linalg.spsolve((K.T @ A0 @ K).tocsc(), K.T @ b0)
    assert np.allclose(K @ d, u, rtol=1e-9, atol=1e-11)
    # periodicity of the solution
    assert np.allclose(u[slaves], u[np.asarray(masters)])
    h = u.copy(); oracle.homogenize(m, h)
    assert np.all(h[slaves] == 0)


@pytest.mark.parametrize("make,measure", [
    (lambda g: g.create_unit_square(4, 3), 4.0), (lambda g: g.create_unit_cube(3, 2, 4), 6.0),
    (lambda g: g.create_rectangle(3, 4, "quadrilateral"), 4.0), (lambda g: g.create_box(2, 3, 2, "hexahedron"), 6.0)])
def test_exterior_facet_integrals_known_answers(oracle, make, measure):
    """Exterior-facet kernels (cpp/assemble_matrix.cpp:271-415, cpp/assemble_vector.cpp:196-240) pinned on
    closed-form answers: int_boundary 1 ds = perimeter / surface area through both the linear form (f = 1) and the
    bilinear form (1^T M 1), P1 and P2; and int_{y=1} x ds = 1/2 on a partial boundary."""
    from dolfinx_mpc_b200 import fem, generators as gen

    mesh = make(gen)
    for degree in ((1, 2) if mesh.cell_type in ("triangle", "tetrahedron") else (1,)):
        V = gen.functionspace(mesh, degree)
        fac = fem.locate_exterior_facets(mesh)
        f = fem.Function(V)
        f.array[:] = 1.0
        e = oracle.OracleMPC.empty(V)
        assert abs(oracle.assemble_vector(fem.source(V, f, 1.0, facets=fac), e).sum() - measure) < 1e-12
        assert abs(oracle.assemble_matrix(fem.mass(V, 1.0, facets=fac), e)[2].sum() - measure) < 1e-12
    mesh = gen.create_unit_square(5, 5)
    V = gen.functionspace(mesh, 1)
    top = fem.locate_exterior_facets(mesh, lambda x: np.isclose(x[1], 1))
    assert len(top) == 5 and np.all(top[:, 1] == 0)
    f = fem.Function(V)
    f.interpolate(lambda x: x[0])
    b = oracle.assemble_vector(fem.source(V, f, 1.0, facets=top), oracle.OracleMPC.empty(V))
    assert abs(b.sum() - 0.5) < 1e-13
